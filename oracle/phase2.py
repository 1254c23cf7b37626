"""Phase-2 (Groth16 MPC) delta update of the H and L queries, restated.

Test infrastructure (see oracle/__init__.py).  Follows `phase2_cli::contribute::<P>`
as called at src/bin/contribute.rs:826-839 and the upstream `batch_mul` (K7) recorded
in SURVEY.md §8a row a10 / Appendix A.5 [UP]:
    h_query[i] *= delta^-1,  l_query[i] *= delta^-1,  delta_g1 *= delta,  delta_g2 *= delta
All four are per-point scalar multiplications with ONE scalar; the oracle performs them
with plain double-and-add.  Only the scaling core is restated: the MPCParameters
container around it (ark-groth16 ProvingKey layout) is not needed to pin the path.
"""
from __future__ import annotations

from . import serialize as ser
from .curves import Curve


def batch_mul(G, pts, k: int):
    return [G.mul(P, k) for P in pts]


def scale_queries(curve: Curve, g1_points_bytes: bytes, delta: int, in_compressed: bool, out_compressed: bool) -> bytes:
    """G1 query vector (h or l) -> the same vector multiplied by delta^-1."""
    dinv = pow(delta, -1, curve.Fr.p)
    pts = ser.points_from_bytes(curve.g1, g1_points_bytes, in_compressed)
    return ser.points_to_bytes(curve.g1, batch_mul(curve.g1, pts, dinv), out_compressed)


def scale_delta(curve: Curve, delta_g1, delta_g2, delta: int):
    return curve.g1.mul(delta_g1, delta), curve.g2.mul(delta_g2, delta)


# ---------------------------------------------------------------------------------------------
# the parameter container and the file-level calls ([UP] phase2::MPCParameters, restated as recalled; csrc/p2.cuh)
# ---------------------------------------------------------------------------------------------
import hashlib
import struct
from dataclasses import dataclass, field


@dataclass
class MPCParameters:
    alpha_g1: object
    beta_g2: object
    gamma_g2: object
    delta_g2: object
    gamma_abc_g1: list
    beta_g1: object
    delta_g1: object
    a_query: list
    b_g1_query: list
    b_g2_query: list
    h_query: list
    l_query: list
    cs_hash: bytes
    contributions: list = field(default_factory=list)          # serialized public keys (always uncompressed)

    def to_bytes(self, curve: Curve, compressed: bool) -> bytes:
        g1, g2 = curve.g1, curve.g2
        p1 = lambda P: ser.point_to_bytes(g1, P, compressed)
        p2 = lambda P: ser.point_to_bytes(g2, P, compressed)
        v1 = lambda v: struct.pack("<Q", len(v)) + ser.points_to_bytes(g1, v, compressed)
        v2 = lambda v: struct.pack("<Q", len(v)) + ser.points_to_bytes(g2, v, compressed)
        return (p1(self.alpha_g1) + p2(self.beta_g2) + p2(self.gamma_g2) + p2(self.delta_g2) + v1(self.gamma_abc_g1) + p1(self.beta_g1)
                + p1(self.delta_g1) + v1(self.a_query) + v1(self.b_g1_query) + v2(self.b_g2_query) + v1(self.h_query) + v1(self.l_query)
                + self.cs_hash + struct.pack(">I", len(self.contributions)) + b"".join(self.contributions))

    @staticmethod
    def from_bytes(curve: Curve, buf: bytes, compressed: bool) -> "MPCParameters":
        g1, g2 = curve.g1, curve.g2
        s1, s2 = ser.point_size(g1, compressed), ser.point_size(g2, compressed)
        o = [0]

        def pt(G, sz):
            P = ser.point_from_bytes(G, buf[o[0]:o[0] + sz], compressed)
            o[0] += sz
            return P

        def vec(G, sz):
            n = struct.unpack("<Q", buf[o[0]:o[0] + 8])[0]
            o[0] += 8
            v = ser.points_from_bytes(G, buf[o[0]:o[0] + n * sz], compressed)
            o[0] += n * sz
            return v
        m = MPCParameters(pt(g1, s1), pt(g2, s2), pt(g2, s2), pt(g2, s2), vec(g1, s1), pt(g1, s1), pt(g1, s1), vec(g1, s1), vec(g1, s1),
                          vec(g2, s2), vec(g1, s1), vec(g1, s1), b"")
        m.cs_hash = buf[o[0]:o[0] + 64]
        o[0] += 64
        n = struct.unpack(">I", buf[o[0]:o[0] + 4])[0]
        o[0] += 4
        csz = 3 * ser.point_size(g1, False) + ser.point_size(g2, False) + 64
        m.contributions = [buf[o[0] + i * csz:o[0] + (i + 1) * csz] for i in range(n)]
        assert o[0] + n * csz == len(buf), "malformed container"
        return m


def transcript_hash(curve: Curve, m: MPCParameters, s, s_delta) -> bytes:
    h = hashlib.blake2b(digest_size=64)
    h.update(m.cs_hash)
    for c in m.contributions:
        h.update(c)
    h.update(ser.point_to_bytes(curve.g1, s, False) + ser.point_to_bytes(curve.g1, s_delta, False))
    return h.digest()


def contribute(curve: Curve, challenge: bytes, rng) -> bytes:
    """phase2_cli::contribute (src/bin/contribute.rs:827-838): challenge container (uncompressed) -> response (compressed)."""
    from .chacha import fp_rand, group_rand
    from .phase1 import hash_to_g2
    m = MPCParameters.from_bytes(curve, challenge, False)
    r = curve.Fr.p
    delta = fp_rand(curve.Fr, rng)
    s = group_rand(curve.g1, rng)
    s_delta = curve.g1.mul(s, delta)
    t = transcript_hash(curve, m, s, s_delta)
    rr = hash_to_g2(curve, t)
    r_delta = curve.g2.mul(rr, delta)
    delta_after = curve.g1.mul(m.delta_g1, delta)
    dinv = pow(delta, -1, r)
    m.h_query = batch_mul(curve.g1, m.h_query, dinv)
    m.l_query = batch_mul(curve.g1, m.l_query, dinv)
    m.delta_g1, m.delta_g2 = delta_after, curve.g2.mul(m.delta_g2, delta)
    pk = (ser.point_to_bytes(curve.g1, delta_after, False) + ser.point_to_bytes(curve.g1, s, False) + ser.point_to_bytes(curve.g1, s_delta, False)
          + ser.point_to_bytes(curve.g2, r_delta, False) + t)
    m.contributions = m.contributions + [pk]
    return m.to_bytes(curve, True)


def verify(curve: Curve, challenge: bytes, response: bytes, rlc=None) -> bytes:
    """phase2_cli::verify (src/bin/contribute.rs:990-1007) with real pairings; returns the new challenge, raises ValueError."""
    import random
    from .pairing import same_ratio
    from .phase1 import hash_to_g2
    g1, g2 = curve.g1, curve.g2
    b = MPCParameters.from_bytes(curve, challenge, False)
    a = MPCParameters.from_bytes(curve, response, True)
    if a.contributions[:-1] != b.contributions or a.cs_hash != b.cs_hash:
        raise ValueError("contribution chain")
    same = lambda x, y, G: all(G.eq(p, q) for p, q in zip(x, y)) and len(x) == len(y)
    if not (same([a.alpha_g1, a.beta_g1] + a.gamma_abc_g1 + a.a_query + a.b_g1_query, [b.alpha_g1, b.beta_g1] + b.gamma_abc_g1 + b.a_query + b.b_g1_query, g1)
            and same([a.beta_g2, a.gamma_g2] + a.b_g2_query, [b.beta_g2, b.gamma_g2] + b.b_g2_query, g2)):
        raise ValueError("an untouched element changed")
    pk = a.contributions[-1]
    s1, s2 = ser.point_size(g1, False), ser.point_size(g2, False)
    delta_after, s, s_delta = (ser.point_from_bytes(g1, pk[i * s1:(i + 1) * s1], False) for i in range(3))
    r_delta = ser.point_from_bytes(g2, pk[3 * s1:3 * s1 + s2], False)
    t = transcript_hash(curve, b, s, s_delta)
    if t != pk[3 * s1 + s2:] or not g1.eq(delta_after, a.delta_g1):
        raise ValueError("public key")
    rr = hash_to_g2(curve, t)
    rnd = rlc or random.Random(1)
    checks = [((s, s_delta), (rr, r_delta)), ((b.delta_g1, a.delta_g1), (rr, r_delta)), ((b.delta_g1, a.delta_g1), (b.delta_g2, a.delta_g2))]
    for before, after in ((b.h_query, a.h_query), (b.l_query, a.l_query)):
        if len(before) != len(after):
            raise ValueError("query length")
        if before:
            rs = [rnd.randrange(1, 1 << 120) for _ in before]
            checks.append(((g1.msm(before, rs), g1.msm(after, rs)), (a.delta_g2, b.delta_g2)))
    for p1, p2 in checks:
        if not same_ratio(curve, p1, p2):
            raise ValueError("same_ratio")
    return a.to_bytes(curve, False)
