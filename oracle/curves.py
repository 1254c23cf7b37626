"""The four pairing curves of the ceremony and their G1 / G2 groups.

Test infrastructure (see oracle/__init__.py).  Restates ark-ec 0.4.2 short
Weierstrass groups (`models/short_weierstrass`) and the curve crates
ark-bls12-377 / ark-bw6-761 / ark-mnt4-753 / ark-mnt6-753 0.4.0
(Cargo.lock:150-151,173-174,282-283,293-294 of the reference).  Constants are
those of SURVEY.md Appendix A.1 (each checked for primality / on-curve / order
in tests/test_oracle_constants.py).

Deliberately simple: affine coordinates, one modular inversion per group
operation, MSB-first double-and-add — a different algorithm from the CUDA
core (Montgomery limbs, Jacobian, GLV windows) so that agreement is evidence.

NOT RECOVERABLE HERE: the arkworks G2 generator constants of MNT4-753 and
MNT6-753.  `Curve.g2.gen` for those two curves is a deterministic order-r point
derived below (`_derive_generator`), flagged by `g2.gen_is_arkworks = False`.
It only matters for `new_challenge` byte parity and the chunk-0 "element 0 is
the generator" check (SURVEY.md §7 hard parts).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from .fields import Fp, ExtField


class Group:
    """y^2 = x^3 + a x + b over `F`; points are None (infinity) or (x, y)."""

    def __init__(self, name, F, a, b, r, gen=None, cofactor=None, gen_is_arkworks=True):
        self.name, self.F, self.a, self.b, self.r = name, F, a, b, r
        self.gen, self._cofactor, self.gen_is_arkworks = gen, cofactor, gen_is_arkworks
        self.a_is_zero = F.is_zero(a)

    @property
    def cofactor(self) -> int:
        """#E(F)/r; an int, or computed on first use from a callable (point counting by trace)."""
        if callable(self._cofactor):
            self._cofactor = self._cofactor()
        return self._cofactor

    @cofactor.setter
    def cofactor(self, v):
        self._cofactor = v

    # -- predicates -----------------------------------------------------------
    def rhs(self, x):
        F = self.F
        return F.add(F.add(F.mul(F.sqr(x), x), F.mul(self.a, x)), self.b)

    def on_curve(self, P) -> bool:
        if P is None:
            return True
        x, y = P
        return self.F.eq(self.F.sqr(y), self.rhs(x))

    def in_subgroup(self, P) -> bool:
        return self.on_curve(P) and self.mul(P, self.r) is None

    # -- group law --------------------------------------------------------------
    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if F.eq(x1, x2):
            if F.eq(y1, y2) and not F.is_zero(y1):
                lam = F.mul(F.add(F.muli(F.sqr(x1), 3), self.a), F.inv(F.muli(y1, 2)))
            else:
                return None
        else:
            lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def double(self, P):
        return self.add(P, P)

    def mul(self, P, k: int):
        """MSB-first double-and-add, as ark-ec `mul_bigint`."""
        if k < 0:
            P, k = self.neg(P), -k
        R = None
        for bit in bin(k)[2:] if k else "":
            R = self.add(R, R)
            if bit == "1":
                R = self.add(R, P)
        return R

    def eq(self, P, Q) -> bool:
        if P is None or Q is None:
            return P is None and Q is None
        return self.F.eq(P[0], Q[0]) and self.F.eq(P[1], Q[1])

    def point_from_x(self, x, greatest: bool):
        """ark-ec `Affine::get_point_from_x_unchecked`: None if rhs is a non-square."""
        F = self.F
        y = F.sqrt(self.rhs(x))
        if y is None:
            return None
        ny = F.neg(y)
        lo, hi = (y, ny) if F.gt(ny, y) else (ny, y)
        return (x, hi if greatest else lo)

    def sum(self, pts):
        R = None
        for P in pts:
            R = self.add(R, P)
        return R

    def msm(self, pts, ks):
        R = None
        for P, k in zip(pts, ks):
            R = self.add(R, self.mul(P, k))
        return R


@dataclass
class Curve:
    name: str            # curveKind string of the coordinator JSON (src/data_structs.rs:123-131)
    cid: int             # id used on the C ABI (include/sso_b200.h)
    Fq: Fp
    Fr: Fp
    g1: Group
    g2: Group
    extra: dict = field(default_factory=dict)


# ---------------------------------------------------------------------------------------------
# constants
# ---------------------------------------------------------------------------------------------
_Q377 = 258664426012969094010652733694893533536393512754914660539884262666720468348340822774968888139573360124440321458177
_R253 = 8444461749428370424248824938781546531375899335154063827935233455917409239041
_X377 = 0x8508c00000000001
_Q761 = 6891450384315732539396789682275657542479668912536150109513790160209623422243491736087683183289411687640864567753786613451161759120554247759349511699125301598951605099378508850372543631423596795951899700429969112842764913119068299
_Q4 = 41898490967918953402344214791240637128170709919953949071783502921025352812571106773058893763790338921418070971888253786114353726529584385201591605722013126468931404347949840543007986327743462853720628051692141265303114721689601
_R4 = 41898490967918953402344214791240637128170709919953949071783502921025352812571106773058893763790338921418070971888458477323173057491593855069696241854796396165721416325350064441470418137846398469611935719059908164220784476160001


def _derive_generator(G: Group, cofactor: int):
    """Deterministic order-r point: smallest x = k + u (k = 0, 1, ...) with a square rhs,
    smaller root, cofactor cleared.  Stand-in where the arkworks constant is unavailable."""
    F = G.F
    k = 0
    while True:
        x = F.from_coeffs((k, 1) + (0,) * (F.deg - 2)) if F.deg > 1 else F.from_int(k)
        P = G.point_from_x(x, False)
        if P is not None:
            Q = G.mul(P, cofactor)
            if Q is not None:
                assert G.mul(Q, G.r) is None
                return Q
        k += 1


def _bls12_377() -> Curve:
    Fq, Fr = Fp(_Q377), Fp(_R253)
    h1 = (_X377 - 1) ** 2 // 3
    g1 = Group("bls12_377.g1", Fq, 0, 1, _R253, gen=(
        81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
        241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030),
        cofactor=h1)
    Fq2 = ExtField(Fq, 2, -5)
    b2 = (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)
    # #E'(Fq2) from the trace of E(Fq): t = q + 1 - h1*r ; CM discriminant -3
    t = _Q377 + 1 - h1 * _R253
    t2 = t * t - 2 * _Q377
    g2 = Group("bls12_377.g2", Fq2, Fq2.zero, b2, _R253, gen=(
        (233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
         140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118),
        (63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
         149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491)))
    g2.cofactor = lambda: _sextic_twist_cofactor(_Q377 ** 2, t2, _R253, g2)
    # GLV: beta = primitive cube root of unity in Fq, lambda = matching root in Fr (phi(x,y) = (beta x, y) = [lambda](x,y))
    return Curve("bls12_377", 0, Fq, Fr, g1, g2)


def _isqrt_exact(n: int) -> int:
    from math import isqrt
    s = isqrt(n)
    assert s * s == n, "not a perfect square"
    return s


def _sextic_twist_cofactor(qk: int, tk: int, r: int, G: Group) -> int:
    """Order of the j=0 twist of E over F_{q^k} that contains G.gen, divided by r.
    Candidates: q^k + 1 - {±t, ±(t ± 3f)/2} with t^2 - 4 q^k = -3 f^2."""
    f = _isqrt_exact((4 * qk - tk * tk) // 3)
    cands = []
    for tr in (tk, -tk, (tk + 3 * f) // 2, (tk - 3 * f) // 2, (-tk + 3 * f) // 2, (-tk - 3 * f) // 2):
        n = qk + 1 - tr
        if n % r == 0:
            cands.append(n)
    good = [n for n in cands if G.mul(_some_point(G), n) is None]
    assert len(good) == 1, (len(cands), len(good))
    return good[0] // r


def _some_point(G: Group, start: int = 1):
    F = G.F
    k = start
    while True:
        x = F.from_coeffs((k, 1) + (0,) * (F.deg - 2)) if F.deg > 1 else F.from_int(k)
        P = G.point_from_x(x, False)
        if P is not None:
            return P
        k += 1


def _bw6_761() -> Curve:
    Fq, Fr = Fp(_Q761), Fp(_Q377)
    g1 = Group("bw6_761.g1", Fq, 0, Fq.from_int(-1), _Q377, gen=(
        6238772257594679368032145693622812838779005809760824733138787810501188623461307351759238099287535516224314149266511977132140828635950940021790489507611754366317801811090811367945064510304504157188661901055903167026722666149426237,
        2101735126520897423911504562215834951148127555913367997162789335052900271653517958562461315794228241561913734371411178226936527683203879553093934185950470971848972085321797958124416462268292467002957525517188485984766314758624099))
    g2 = Group("bw6_761.g2", Fq, 0, 4, _Q377, gen=(
        6445332910596979336035888152774071626898886139774101364933948236926875073754470830732273879639675437155036544153105017729592600560631678554299562762294743927912429096636156401171909259073181112518725201388196280039960074422214428,
        562923658089539719386922163444547387757586534741080263946953401595155211934630598999300396317104182598044793758153214972605680357108252243146746187917218885078195819486220416605630144001533548163105316661692978285266378674355041))
    # cofactors: #E = r*h with |q + 1 - #E| <= 2 sqrt(q); scan the few admissible h
    g1.cofactor = lambda: _scan_cofactor(g1, _Q761)
    g2.cofactor = lambda: _scan_cofactor(g2, _Q761)
    return Curve("bw6_761", 1, Fq, Fr, g1, g2)


def _scan_cofactor(G: Group, q: int) -> int:
    from math import isqrt
    r = G.r
    lo = (q + 1 - 2 * isqrt(q) - 2) // r
    hi = (q + 1 + 2 * isqrt(q) + 2) // r + 1
    bases = [G.mul(_some_point(G, s), r) for s in (1, 100, 10000, 1000000)]
    good = [h for h in range(lo, hi + 1) if all(G.mul(B, h) is None for B in bases)]
    assert len(good) == 1, good
    return good[0]


def _mnt4_753() -> Curve:
    Fq, Fr = Fp(_Q4), Fp(_R4)
    a = 2
    b = 28798803903456388891410036793299405764940372360099938340752576406393880372126970068421383312482853541572780087363938442377933706865252053507077543420534380486492786626556269083255657125025963825610840222568694137138741554679540
    g1 = Group("mnt4_753.g1", Fq, a, b, _R4, gen=(
        7790163481385331313124631546957228376128961350185262705123068027727518350362064426002432450801002268747950550964579198552865939244360469674540925037890082678099826733417900510086646711680891516503232107232083181010099241949569,
        6913648190367314284606685101150155872986263667483624713540251048208073654617802840433842931301128643140890502238233930290161632176167186761333725658542781350626799660920481723757654531036893265359076440986158843531053720994648),
        cofactor=1)
    Fq2 = ExtField(Fq, 2, 13)
    g2 = Group("mnt4_753.g2", Fq2, (a * 13 % _Q4, 0), (0, b * 13 % _Q4), _R4, gen_is_arkworks=False)
    t = _Q4 + 1 - _R4
    n2 = _Q4 ** 2 + 1 + (t * t - 2 * _Q4)            # quadratic twist of E over Fq2
    assert n2 % _R4 == 0
    g2.cofactor = n2 // _R4
    g2.gen = _derive_generator(g2, g2.cofactor)
    return Curve("mnt4_753", 2, Fq, Fr, g1, g2)


def _mnt6_753() -> Curve:
    Fq, Fr = Fp(_R4), Fp(_Q4)
    a = 11
    b = 11625908999541321152027340224010374716841167701783584648338908235410859267060079819722747939267925389062611062156601938166010098747920378738927832658133625454260115409075816187555055859490253375704728027944315501122723426879114
    g1 = Group("mnt6_753.g1", Fq, a, b, _Q4, gen=(
        16364236387491689444759057944334173579070747473738339749093487337644739228935268157504218078126401066954815152892688541654726829424326599038522503517302466226143788988217410842672857564665527806044250003808514184274233938437290,
        4510127914410645922431074687553594593336087066778984214797709122300210966076979927285161950203037801392624582544098750667549188549761032654706830225743998064330900301346566408501390638273322467173741629353517809979540986561128),
        cofactor=1)
    Fq3 = ExtField(Fq, 3, 11)
    g2 = Group("mnt6_753.g2", Fq3, (0, 0, a), (b * 11 % _R4, 0, 0), _Q4, gen_is_arkworks=False)
    q = _R4
    t = q + 1 - _Q4
    n3 = q ** 3 + 1 + (t ** 3 - 3 * q * t)           # quadratic twist of E over Fq3
    assert n3 % _Q4 == 0
    g2.cofactor = n3 // _Q4
    g2.gen = _derive_generator(g2, g2.cofactor)
    return Curve("mnt6_753", 3, Fq, Fr, g1, g2)


_BUILDERS = {"bls12_377": _bls12_377, "bw6_761": _bw6_761, "mnt4_753": _mnt4_753, "mnt6_753": _mnt6_753}
_ALIASES = {"bw6": "bw6_761"}            # curveKind strings accepted by the operator (src/bin/new_setup.rs:53-54)
_CACHE: dict = {}

CURVE_NAMES = tuple(_BUILDERS)


def get_curve(name: str) -> Curve:
    name = _ALIASES.get(name, name)
    if name not in _CACHE:
        _CACHE[name] = _BUILDERS[name]()
    return _CACHE[name]
