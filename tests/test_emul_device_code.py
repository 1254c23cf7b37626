"""Runs the per-thread bodies of the CUDA kernels on the CPU (tests/emul/emul.cpp: the same
.cuh sources compiled by g++ with an instruction-level emulation of the PTX carry flag) and
compares them with the oracle.  This is how the device algorithms are exercised in the
GPU-less container; the bit-exact GPU runs are in test_gpu_parity.py."""
import ctypes
import os
import random
import subprocess

import pytest

from conftest import ROOT
from oracle import serialize as ser, synth
from oracle.curves import CURVE_NAMES, get_curve

EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libsso_emul.so")
    srcs = [os.path.join(EMUL_DIR, "emul.cpp")] + [os.path.join(ROOT, "snark-setup-operator_b200", "csrc", f)
                                                    for f in os.listdir(os.path.join(ROOT, "snark-setup-operator_b200", "csrc"))
                                                    if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so, os.path.join(EMUL_DIR, "emul.cpp")])
    return ctypes.CDLL(so)


def words(v, nwords):
    return (ctypes.c_uint32 * nwords)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(nwords)])


FIELDS = {0: ("bls12_377", "Fr"), 1: ("bls12_377", "Fq"), 2: ("bw6_761", "Fq"), 3: ("mnt4_753", "Fq"), 4: ("mnt6_753", "Fq"),
          5: ("bls12_377", "g2"), 6: ("mnt4_753", "g2"), 7: ("mnt6_753", "g2"),
          # the warp-cooperative extension fields (coop.cuh): one coefficient per lane, lanes emulated by lockstep threads
          8: ("bls12_377", "g2"), 9: ("mnt4_753", "g2"), 10: ("mnt6_753", "g2")}


def _field(fid):
    name, which = FIELDS[fid]
    c = get_curve(name)
    return c.g2.F if which == "g2" else getattr(c, which)


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_field_ops(emul, fid):
    F = _field(fid)
    rnd = random.Random(fid)

    def relem():
        return rnd.randrange(F.p) if F.deg == 1 else tuple(rnd.randrange(F.p) for _ in range(F.deg))

    ops = {0: lambda a, b: F.mul(a, b), 1: lambda a, b: F.add(a, b), 2: lambda a, b: F.sub(a, b), 3: lambda a, b: F.sqr(a),
           4: lambda a, b: F.neg(a), 5: lambda a, b: F.inv(a), 6: lambda a, b: F.sqr(a)}     # 6 = dedicated squaring
    cases = [(relem(), relem()) for _ in range(6)]
    top = F.p - 1 if F.deg == 1 else (F.p - 1,) * F.deg
    cases += [(F.zero, relem()), (top, top), (F.one, top)]
    if F.deg > 1:
        # coefficient patterns for the unreduced-operand paths of the extension multiplications (zero / p - 1 coefficients)
        pm1 = F.p - 1
        cases += [(tuple(pm1 if (m >> i) & 1 else 0 for i in range(F.deg)), tuple(pm1 if (m >> (i + 1)) & 1 else 1 for i in range(F.deg)))
                  for m in range(1, 1 << F.deg)]
        cases += [(tuple(rnd.randrange(F.p) if i != z else 0 for i in range(F.deg)), relem()) for z in range(F.deg)]
    if F.deg == 1:
        # carry-heavy residues for the dedicated squaring: runs of 0xffffffff / 0x00000000 limbs, single bits, p - small
        bits = F.p.bit_length()
        patt = [(1 << (bits - 1)) - 1, (1 << (bits - 2)) + 1, int("ffffffff00000000" * 12, 16) % F.p, int("00000000ffffffff" * 12, 16) % F.p,
                (1 << 32) - 1, 1 << 32, F.p - 2, F.p // 2, F.p // 2 + 1, (1 << (32 * (bits // 32))) - 1]
        patt += [rnd.randrange(F.p) | int("ffffffff" * (bits // 64), 16) for _ in range(6)]
        cases += [(x % F.p, relem()) for x in patt]
    for a, b in cases:
        ab, bb = ser.field_to_bytes(F, a), ser.field_to_bytes(F, b)
        for op, fn in ops.items():
            if op == 5 and F.is_zero(a):
                continue
            out = ctypes.create_string_buffer(len(ab))
            assert emul.emul_field_op(fid, op, ab, bb, out) == 0
            assert out.raw == ser.field_to_bytes(F, fn(a, b)), (fid, op)
        if fid < 8:
            out = ctypes.create_string_buffer(len(ab))
            assert emul.emul_field_op(fid, 7, ab, ab, out) == (1 if F.gt(a, F.neg(a)) else 0)


@pytest.mark.parametrize("name", CURVE_NAMES)
@pytest.mark.parametrize("gi", [0, 1])
def test_batch_exp_and_reencode(emul, name, gi):
    c = get_curve(name)
    G = (c.g1, c.g2)[gi]
    if c.Fq.bits > 400 and gi == 1 and name != "bw6_761":
        n = 3
    else:
        n = 4
    rnd = random.Random(7 * c.cid + gi)
    Lr = (c.Fr.bits + 31) // 32
    key = synth.contributor_key(c)
    pts = [G.mul(G.gen, rnd.randrange(1, G.r)) for _ in range(n - 1)] + [None]
    buf = ser.points_to_bytes(G, pts, False)
    first = 1000003
    want_pts = [G.mul(P, key.beta * pow(key.tau, first + j, c.Fr.p) % c.Fr.p) for j, P in enumerate(pts)]
    want = ser.points_to_bytes(G, want_pts, True)
    out = ctypes.create_string_buffer(len(want))
    st = (ctypes.c_uint32 * 3)()
    assert emul.emul_batch_exp(c.cid, gi, buf, 0, n, words(key.tau, Lr), words(key.beta, Lr), ctypes.c_uint64(first), 0, 1,
                               out, 1, st) == 0
    assert out.raw == want
    assert list(st)[:2] == [4, n - 1]                       # the point at infinity is reported under CHECK_NONZERO
    # shared-scalar mode (phase-2 batch_mul), uncompressed output
    want2 = ser.points_to_bytes(G, [G.mul(P, key.alpha) for P in pts], False)
    out2 = ctypes.create_string_buffer(len(want2))
    assert emul.emul_batch_exp(c.cid, gi, buf, 0, n, words(1, Lr), words(key.alpha, Lr), ctypes.c_uint64(0), 1, 0, out2, 0,
                               st) == 0
    assert out2.raw == want2 and list(st)[:2] == [0, 0]
    # decompression + full checks incl. subgroup membership
    out3 = ctypes.create_string_buffer(len(buf))
    assert emul.emul_reencode(c.cid, gi, want, 1, n - 1, out3, 0, 2, 1, st) == 0
    assert out3.raw[:len(buf) - ser.point_size(G, False)] == ser.points_to_bytes(G, want_pts[:-1], False)
    assert list(st)[:2] == [0, 0]


@pytest.mark.parametrize("body", [1, 2])                    # 1 = one thread per element, 2 = warp-cooperative (coop.cuh)
def test_bls12_g2_four_way_decomposition_edge_scalars(emul, body):
    """The G2 ladder of BLS12-377 splits the scalar into base-x digits (x = curve parameter) and uses psi, psi^2, psi^3:
    scalars on the digit boundaries, and a small-order base point (Jacobian-table path), must still match [k]P."""
    c = get_curve("bls12_377")
    G = c.g2
    x = 0x8508c00000000001
    r = c.Fr.p
    Lr = (c.Fr.bits + 31) // 32
    pts = [G.mul(G.gen, 0x1234567), G.mul(G.gen, r - 5)]
    buf = ser.points_to_bytes(G, pts, False)
    st = (ctypes.c_uint32 * 3)()
    ks = [1, 2, 8, 9, 15, 16, x - 1, x, x + 1, x * x - 1, x * x, x ** 3 - 1, x ** 3, x ** 3 + x * x + x + 1, r - 1, r - 2,
          (x - 1) * (1 + x + x * x), 0x8888888888888888, (x - 1) + (x - 1) * x + (x - 1) * x * x + ((r - 1) // x ** 3) * x ** 3]
    for k in ks:
        k %= r
        if k == 0:
            continue
        want = ser.points_to_bytes(G, [G.mul(P, k) for P in pts], False)
        out = ctypes.create_string_buffer(len(want))
        assert emul.emul_batch_exp(c.cid, body, buf, 0, 2, words(1, Lr), words(k, Lr), ctypes.c_uint64(0), 1, 0, out, 0, st) == 0
        assert out.raw == want, hex(k)


@pytest.mark.parametrize("body", [1, 2])
@pytest.mark.parametrize("name", ["mnt4_753", "mnt6_753"])
def test_mnt_g2_two_way_decomposition_edge_scalars(emul, name, body):
    """The G2 ladders of the MNT curves split the scalar as k = k0 + k1 mu (mu = |t - 1|, the eigenvalue of the Frobenius
    endomorphism psi up to sign; Barrett quotient with corrections) and add psi-images of the table entries: scalars on the
    digit and quotient boundaries, and a base point of the full group order, must still match [k]P."""
    c = get_curve(name)
    G = c.g2
    r = c.Fr.p
    mu = abs(c.Fq.p - r)                                    # |t - 1| = |q - r|
    Lr = (c.Fr.bits + 31) // 32
    pts = [G.mul(G.gen, 0x1234567), G.mul(G.gen, r - 5)]
    buf = ser.points_to_bytes(G, pts, False)
    st = (ctypes.c_uint32 * 3)()
    ks = [1, 2, 7, 8, 9, 15, 16, mu - 1, mu, mu + 1, 2 * mu - 1, 2 * mu, 3 * mu + 5, (r // mu) * mu, (r // mu) * mu - 1, r - 1, r - 2, r // 2,
          0x8888888888888888 * mu + 0x7777777777777777, (mu - 1) + (mu - 1) * mu if (mu - 1) + (mu - 1) * mu < r else r - 3]
    for k in ks:
        k %= r
        if k == 0:
            continue
        want = ser.points_to_bytes(G, [G.mul(P, k) for P in pts], False)
        out = ctypes.create_string_buffer(len(want))
        assert emul.emul_batch_exp(c.cid, body, buf, 0, 2, words(1, Lr), words(k, Lr), ctypes.c_uint64(0), 1, 0, out, 0, st) == 0
        assert out.raw == want, hex(k)


@pytest.mark.parametrize("name,gi", [("bls12_377", 0), ("bls12_377", 1), ("bw6_761", 0), ("bw6_761", 1), ("mnt4_753", 1), ("mnt6_753", 1)])
def test_endomorphism_subgroup_tests_agree_with_order_check(emul, name, gi):
    """Membership is tested on the device in endomorphism form where the curve has one — BLS12-377: phi(P) = [-x^2]P (G1) /
    psi(P) = [x]P (G2); BW6-761 G1: [x + 1]P + [x^3 - x^2 + 1]phi(P) = O (G2 keeps [r]P); MNT4/6-753 G2: psi(P) = [t - 1]P —; the verdict must be the
    reference's [r]P == O on every kind of on-curve point: random curve points, pure cofactor-torsion points [r]P, points of
    tiny order, and subgroup points plus a cofactor-torsion component."""
    c = get_curve(name)
    G = c.g1 if gi == 0 else c.g2
    from oracle.curves import _some_point
    rogue = [_some_point(G, s) for s in (11, 12, 13)]
    torsion = [G.mul(P, G.r) for P in rogue[:2]]                      # order divides the cofactor
    small = [G.mul(rogue[0], (G.cofactor * G.r) // ell) for ell in (2, 3, 7) if G.cofactor % ell == 0]   # tiny orders
    mixed = [G.add(G.mul(G.gen, 77), torsion[0])]
    good = [G.mul(G.gen, k) for k in (1, 2, 0x123456789abcdef, G.r - 1)]
    cases = [(P, False) for P in rogue + torsion + mixed + small] + [(P, True) for P in good]
    st = (ctypes.c_uint32 * 3)()
    sz = ser.point_size(G, False)
    out = ctypes.create_string_buffer(sz)
    for P, want_ok in cases:
        if P is None:
            continue
        assert G.on_curve(P) and G.in_subgroup(P) == want_ok
        emul.emul_reencode(c.cid, gi, ser.point_to_bytes(G, P, False), 0, 1, out, 0, 2, 1, st)
        assert (st[0] == 0) == want_ok, (name, gi, want_ok, st[0])
        if not want_ok:
            assert st[0] == 5


def test_reencode_rejects_bad_points(emul):
    c = get_curve("bls12_377")
    G = c.g1
    from oracle.curves import _some_point
    rogue = _some_point(G, 11)                           # on the curve, not in the r-torsion
    st = (ctypes.c_uint32 * 3)()
    out = ctypes.create_string_buffer(96)
    emul.emul_reencode(c.cid, 0, ser.point_to_bytes(G, rogue, True), 1, 1, out, 0, 2, 1, st)
    assert list(st)[:2] == [5, 0]
    emul.emul_reencode(c.cid, 0, ser.point_to_bytes(G, rogue, True), 1, 1, out, 0, 2, 0, st)
    assert list(st)[:2] == [0, 0] and out.raw == ser.point_to_bytes(G, rogue, False)
    bad = bytearray(ser.point_to_bytes(G, G.gen, True)); bad[-1] |= 0xC0
    emul.emul_reencode(c.cid, 0, bytes(bad), 1, 1, out, 0, 2, 1, st)
    assert st[0] == 2
    noncanon = (c.Fq.p).to_bytes(48, "little")
    emul.emul_reencode(c.cid, 0, noncanon, 1, 1, out, 0, 2, 1, st)
    assert st[0] == 1
    # x with no y on the curve
    x = 0
    while G.point_from_x(x, False) is not None:
        x += 1
    emul.emul_reencode(c.cid, 0, ser.field_to_bytes(c.Fq, x), 1, 1, out, 0, 2, 1, st)
    assert st[0] == 3
    # off-curve uncompressed point
    offc = ser.field_to_bytes(c.Fq, 5) + ser.field_to_bytes(c.Fq, 7)
    out2 = ctypes.create_string_buffer(48)
    emul.emul_reencode(c.cid, 0, offc, 0, 1, out2, 1, 2, 0, st)
    assert st[0] == 3


@pytest.mark.parametrize("name", ["bls12_377", "bw6_761", "mnt4_753"])
def test_edge_scalars(emul, name):
    """Scalars that stress the window recoding and the GLV split: 0, 1, r-1, powers of two around
    sqrt(r), all-ones patterns; and P = generator so that table entries collide with the accumulator."""
    c = get_curve(name)
    r = c.Fr.p
    Lr = (c.Fr.bits + 31) // 32
    half = c.Fr.bits // 2
    ks = [0, 1, 2, 7, 8, 9, 15, 16, r - 1, r - 2, (r - 1) // 2, 1 << half, (1 << half) - 1, (1 << (half + 1)) + 1,
          (1 << (c.Fr.bits - 1)) - 1, int("8" * (c.Fr.bits // 4 - 1), 16), int("7" * (c.Fr.bits // 4 - 1), 16)]
    gi = 0
    G = c.g1
    buf = ser.points_to_bytes(G, [G.gen], False)
    st = (ctypes.c_uint32 * 3)()
    for k in ks:
        want = ser.points_to_bytes(G, [G.mul(G.gen, k)], True)
        out = ctypes.create_string_buffer(len(want))
        assert emul.emul_batch_exp(c.cid, gi, buf, 0, 1, words(1, Lr), words(k, Lr), ctypes.c_uint64(0), 1, 0, out, 1, st) == 0
        assert out.raw == want, hex(k)


@pytest.mark.parametrize("name,gi", [("bls12_377", 0), ("bls12_377", 1), ("mnt6_753", 0)])
def test_multi_vector_launch(emul, name, gi):
    """Two vectors with different coefficient slots and scalar rules in one batch (the phase-1 chunk
    launches tauG1|alphaG1|betaG1 and tauG2|betaG2 this way)."""
    c = get_curve(name)
    G = (c.g1, c.g2)[gi]
    r = c.Fr.p
    Lr = (c.Fr.bits + 31) // 32
    rnd = random.Random(99)
    key = synth.contributor_key(c)
    v0 = [G.mul(G.gen, rnd.randrange(1, r)) for _ in range(3)]
    v1 = [G.mul(G.gen, rnd.randrange(1, r))]
    first = 77
    want0 = ser.points_to_bytes(G, [G.mul(P, key.alpha * pow(key.tau, first + j, r) % r) for j, P in enumerate(v0)], True)
    want1 = ser.points_to_bytes(G, [G.mul(v1[0], key.beta)], True)
    out0, out1 = ctypes.create_string_buffer(len(want0)), ctypes.create_string_buffer(len(want1))
    st = (ctypes.c_uint32 * 3)()
    assert emul.emul_batch_exp2(c.cid, gi, ser.points_to_bytes(G, v0, False), 3, ser.points_to_bytes(G, v1, False), 1,
                                words(key.tau, Lr), words(key.alpha, Lr), words(key.beta, Lr), ctypes.c_uint64(first), out0, out1,
                                st) == 0
    assert out0.raw == want0 and out1.raw == want1 and st[0] == 0


@pytest.mark.parametrize("name,gi,wb", [("bls12_377", 0, 4), ("bls12_377", 1, 7), ("bls12_377", 0, 9), ("bw6_761", 0, 5), ("mnt4_753", 0, 6)])
def test_power_pairs_msm(emul, name, gi, wb):
    """Pippenger pipeline (keys, buckets, segmented fold, window sums, Horner) against the oracle's
    plain sum r_i P_i with the same ChaCha20-derived scalars."""
    from oracle import phase1
    c = get_curve(name)
    G = (c.g1, c.g2)[gi]
    rnd = random.Random(5 + gi)
    n = 9
    pts = [G.mul(G.gen, rnd.randrange(1, G.r)) for _ in range(n)]
    pts[3] = None                                           # infinity inside the vector
    seed = bytes(range(100, 132))
    rs = phase1.rlc_scalars(c, seed, n - 1)
    a, b = phase1.power_pairs_with(G, pts, rs)
    want = ser.point_to_bytes(G, a, False) + ser.point_to_bytes(G, b, False)
    out = ctypes.create_string_buffer(len(want))
    st = (ctypes.c_uint32 * 3)()
    seed_words = (ctypes.c_uint32 * 8).from_buffer_copy(seed)
    assert emul.emul_power_pairs(c.cid, gi, ser.points_to_bytes(G, pts, True), 1, n, seed_words, wb, out, None, st) == 0
    assert out.raw == want


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_same_ratio_pairing(emul, name):
    """Device Tate pairing (lane-per-coefficient Fq^k arithmetic, Miller loop, final exponentiation)
    against the oracle's verdicts: equal ratios accept, unequal reject, infinity handled."""
    from oracle import pairing
    c = get_curve(name)
    g1, g2 = c.g1, c.g2
    P = g1.mul(g1.gen, 0x1234567)
    Q = g2.mul(g2.gen, 0x7654321)
    x = 0xABCDEF0123456789
    cases = [((P, g1.mul(P, x)), (Q, g2.mul(Q, x)), True),
             ((P, g1.mul(P, x)), (Q, g2.mul(Q, x + 1)), False),
             ((g1.gen, g1.gen), (g2.gen, g2.gen), True)]
    buf = b""
    for (a, b), (cc, d), _ in cases:
        buf += ser.point_to_bytes(g1, a, False) + ser.point_to_bytes(g1, b, False)
        buf += ser.point_to_bytes(g2, cc, False) + ser.point_to_bytes(g2, d, False)
    verdicts = (ctypes.c_uint32 * len(cases))()
    assert emul.emul_same_ratio(c.cid, buf, len(cases), verdicts) == 0
    assert [int(v) for v in verdicts] == [1 if w else 0 for _, _, w in cases]
    if name == "bls12_377":
        for (a, b), (cc, d), want in cases[:2]:
            assert pairing.same_ratio(c, (a, b), (cc, d)) == want


@pytest.mark.parametrize("name", ["bls12_377", "mnt4_753"])
def test_key_generation_matches_oracle(emul, name):
    """Device samplers (ChaCha20 stream, Fr::rand, G::rand with cofactor clearing, hash_to_g2) reproduce
    the oracle's Phase1::key_generation byte for byte."""
    import hashlib
    from oracle import phase1
    from oracle.chacha import ChaChaRng
    c = get_curve(name)
    seed = synth.SEED_CONTRIB
    digest = phase1.blank_hash()
    pub, key = phase1.key_generation(c, ChaChaRng(seed), digest)
    Lr = (c.Fr.bits + 31) // 32
    g1u, g2u = ser.point_size(c.g1, False), ser.point_size(c.g2, False)
    scal = (ctypes.c_uint32 * (3 * Lr))()
    g1_out = ctypes.create_string_buffer(6 * g1u)
    assert emul.emul_keygen_g1(c.cid, (ctypes.c_uint32 * 8).from_buffer_copy(seed), 3, scal, g1_out) == 0
    got = [sum(scal[i * Lr + w] << (32 * w) for w in range(Lr)) for i in range(3)]
    assert got == [key.tau, key.alpha, key.beta]
    want_pk = pub.to_bytes(c)
    assert g1_out.raw == want_pk[:6 * g1u]
    seeds = b""
    for i in range(3):
        h = hashlib.blake2b(digest_size=64)
        h.update(bytes([i])); h.update(digest); h.update(g1_out.raw[2 * i * g1u:(2 * i + 2) * g1u])
        seeds += h.digest()[:32]
    g2_s = ctypes.create_string_buffer(3 * g2u)
    g2_sx = ctypes.create_string_buffer(3 * g2u)
    assert emul.emul_hash_to_g2(c.cid, 3, (ctypes.c_uint32 * 24).from_buffer_copy(seeds), scal, g2_s, g2_sx) == 0
    assert g2_sx.raw == want_pk[6 * g1u:]
    assert g2_s.raw[:g2u] == ser.point_to_bytes(c.g2, phase1.compute_g2_s(c, digest, pub.tau_g1[0], pub.tau_g1[1], 0), False)


def test_small_order_base_point(emul):
    """A base point whose multiples hit the identity (order 2: (-1, 0) on y^2 = x^3 + 1) makes table entries with
    Z = 0; the affine-table path must fall back to the Jacobian table for that thread only, and the other threads of
    the block (sharing the batched inversion) must be unaffected."""
    c = get_curve("bls12_377")
    G = c.g1
    p2 = (c.Fq.p - 1, 0)
    assert G.on_curve(p2) and G.add(p2, p2) is None
    Lr = (c.Fr.bits + 31) // 32
    pts = [G.mul(G.gen, 5), p2, G.mul(G.gen, 7), None]
    buf = ser.points_to_bytes(G, pts, False)
    st = (ctypes.c_uint32 * 3)()
    for k in (1, 2, 12345, 12346):
        want = ser.points_to_bytes(G, [G.mul(P, k) for P in pts], True)
        out = ctypes.create_string_buffer(len(want))
        assert emul.emul_batch_exp(c.cid, 0, buf, 0, len(pts), words(1, Lr), words(k, Lr), ctypes.c_uint64(0), 1, 0, out, 1, st) == 0
        assert out.raw == want, k


@pytest.mark.parametrize("name", ["bls12_377", "mnt4_753", "mnt6_753"])
def test_cooperative_g2_bodies(emul, name):
    """coop.cuh: an Fq2 / Fq3 element is held by two / three adjacent lanes (one coefficient each); the lanes of a group are
    emulated by lockstep host threads and their exchanges by a shared slot array.  Same inputs, same bytes as the
    one-thread-per-element bodies: tau powers with a coefficient, the shared-scalar mode, an infinity element under
    CHECK_NONZERO, the on-curve check under CHECK_FULL (accept and reject), more points than one group."""
    c = get_curve(name)
    G = c.g2
    n = 6 if name == "bls12_377" else 4
    rnd = random.Random(31 * c.cid)
    Lr = (c.Fr.bits + 31) // 32
    key = synth.contributor_key(c)
    pts = [G.mul(G.gen, rnd.randrange(1, G.r)) for _ in range(n - 1)] + [None]
    buf = ser.points_to_bytes(G, pts, False)
    first = (1 << 24) + 12345
    want = ser.points_to_bytes(G, [G.mul(P, key.alpha * pow(key.tau, first + j, c.Fr.p) % c.Fr.p) for j, P in enumerate(pts)], True)
    out = ctypes.create_string_buffer(len(want))
    st = (ctypes.c_uint32 * 3)()
    assert emul.emul_batch_exp(c.cid, 2, buf, 0, n, words(key.tau, Lr), words(key.alpha, Lr), ctypes.c_uint64(first), 0, 1, out, 1, st) == 0
    assert out.raw == want
    assert list(st)[:2] == [4, n - 1]
    # CHECK_FULL accepts curve points; a point off the curve (y + 1) is reported at its index
    want2 = ser.points_to_bytes(G, [G.mul(P, key.beta) for P in pts[:-1]], False)
    out2 = ctypes.create_string_buffer(len(want2))
    assert emul.emul_batch_exp(c.cid, 2, buf, 0, n - 1, words(1, Lr), words(key.beta, Lr), ctypes.c_uint64(0), 1, 2, out2, 0, st) == 0
    assert out2.raw == want2 and list(st)[:2] == [0, 0]
    F = G.F
    bad = list(pts[:-1])
    bad[1] = (bad[1][0], F.add(bad[1][1], F.one))
    assert emul.emul_batch_exp(c.cid, 2, ser.points_to_bytes(G, bad, False), 0, n - 1, words(1, Lr), words(key.beta, Lr),
                               ctypes.c_uint64(0), 1, 2, out2, 0, st) == 0
    assert list(st)[:2] == [3, 1]                            # ST_NOT_ON_CURVE at element 1
    # compressed input is not taken by the cooperative bodies (the host selects the other kernel)
    assert emul.emul_batch_exp(c.cid, 2, want, 1, n, words(1, Lr), words(key.beta, Lr), ctypes.c_uint64(0), 1, 0, out2, 0, st) == -2


@pytest.mark.parametrize("fid", [1, 2, 3, 4])
def test_unreduced_operands_at_their_bounds(emul, fid):
    """fp.cuh "lazy operands": mont_mul returns the canonical a b R^-1 mod p for operands below A p and B p whenever A B p <= R,
    the dedicated squaring for A^2 p <= R, reduce_small for any y < 16 p, mad_small_lazy is the integer c + k x.  Probed at the
    LARGEST magnitudes the extension-field products and the point formulas produce: operands of the form
    (A - 1) p + (p - 1), all-ones limb patterns below the bound, and random ones."""
    F = _field(fid)
    p = F.p
    L = (F.bits + 31) // 32
    R = 1 << (32 * L)
    Rinv = pow(R, -1, p)
    rnd = random.Random(100 + fid)
    ratio = R // p                                            # 152 (377 bits), 225 (761), 37054 (753)

    def arr(v):
        assert 0 <= v < R
        return (ctypes.c_uint32 * L)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(L)])

    def val(a):
        return sum(int(a[i]) << (32 * i) for i in range(L))

    out = (ctypes.c_uint32 * L)()
    # (A, B) pairs used by the code, largest products first: Fq2 squaring on BLS12-377 (6 p, 23 p), cooperative / plain MNT4
    # squaring (58 p, 406 p), MNT6 G1 doubling (14 p, 14 p), point formulas (3 p, 3 p), (4 p, 2 p), ...
    pairs = {152: [(6, 23), (4, 6), (3, 3), (2, 12)], 225: [(3, 3), (4, 4), (4, 2), (15, 15)],
             37054: [(58, 406), (29, 754 // 29), (14, 14), (125, 2), (192, 192)]}[ratio]
    for A, B in pairs:
        assert A * B <= ratio
        cases = [(A * p - 1, B * p - 1), ((A - 1) * p + rnd.randrange(p), (B - 1) * p + rnd.randrange(p)), (A * p - 1, 1), (0, B * p - 1)]
        cases += [(min(A * p - 1, (1 << (A * p).bit_length() - 1) - 1), min(B * p - 1, (1 << (B * p).bit_length() - 1) - 1))]
        for a, b in cases:
            assert emul.emul_lazy_probe(fid, 0, arr(a), arr(b), 0, out) == 0
            assert val(out) == a * b * Rinv % p, (fid, A, B)
    # dedicated squaring (12-limb fields use it in the point formulas): arguments up to 3 p (3 X^2) and 4 p
    if L <= 12:
        for A in (2, 3, 4):
            for a in (A * p - 1, (A - 1) * p + rnd.randrange(p)):
                assert emul.emul_lazy_probe(fid, 3, arr(a), arr(0), 0, out) == 0
                assert val(out) == a * a * Rinv % p
    # reduce_small over [0, 16 p)
    for y in [0, p - 1, p, p + 1, 2 * p - 1, 8 * p - 1, 15 * p + (p - 1), 16 * p - 1] + [rnd.randrange(16 * p) for _ in range(20)]:
        assert emul.emul_lazy_probe(fid, 1, arr(y), arr(0), 0, out) == 0
        assert val(out) == y % p, hex(y)
    # mad_small_lazy: c + k x as integers
    for k in (0, 1, 5, 13, 14, 26):
        for c, x in ((p - 1, p - 1), (rnd.randrange(3 * p), rnd.randrange(p)), (0, p - 1)):
            if c + k * x >= R:
                continue
            assert emul.emul_lazy_probe(fid, 2, arr(c), arr(x), k, out) == 0
            assert val(out) == c + k * x
