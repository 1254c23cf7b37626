"""Parity of libsso_b200.so (through the C ABI, on a real GPU) with the oracle:
bit-exact against the committed golden chunks, against live oracle runs on small seeded inputs,
and through size-independent properties at larger sizes."""
import ctypes
import os
import random

import pytest

import snark_setup_operator_b200 as sso
from oracle import phase1, serialize as ser, synth
from oracle.curves import CURVE_NAMES, get_curve
from oracle.params import Phase1Params

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev_bytes(b):
    return torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()


def host_bytes(t):
    return bytes(t.cpu().numpy().tobytes())


FIELD_IDS = {0: ("bls12_377", "Fr"), 1: ("bls12_377", "Fq"), 2: ("bw6_761", "Fq"), 3: ("mnt4_753", "Fq"), 4: ("mnt6_753", "Fq")}


@pytest.mark.parametrize("fid", sorted(FIELD_IDS))
def test_field_mul_bit_exact(fid):
    name, which = FIELD_IDS[fid]
    F = getattr(get_curve(name), which)
    rnd = random.Random(100 + fid)
    n = 2048
    a = [rnd.randrange(F.p) for _ in range(n)]
    b = [rnd.randrange(F.p) for _ in range(n)]
    a[:4] = [0, 1, F.p - 1, F.p - 1]
    b[:4] = [5, F.p - 1, F.p - 1, 2]
    A = b"".join(ser.field_to_bytes(F, x) for x in a)
    B = b"".join(ser.field_to_bytes(F, x) for x in b)
    out = ctypes.create_string_buffer(len(A))
    sso._lib.call("sso_test_field_mul", fid, A, B, out, n, 0)
    want = b"".join(ser.field_to_bytes(F, x * y % F.p) for x, y in zip(a, b))
    assert out.raw == want


@pytest.mark.parametrize("name", ["bls12_377", "mnt4_753", "mnt6_753"])
def test_cooperative_and_single_thread_g2_bodies_agree(name, monkeypatch):
    """csrc/coop.cuh: the G2 batch_exp bodies with two / three lanes per Fq2 / Fq3 element against the one-thread-per-element
    bodies (SSO_COOP_G2 = 1 / 0 forces either; the default is per curve) and against the oracle: a ragged count that leaves the
    last group of lanes, the last warp and the last block partly empty, an infinity element, CHECK_FULL on curve points, and a
    chunk contribution through both."""
    c = get_curve(name)
    G = c.g2
    rnd = random.Random(17 * c.cid)
    key = synth.contributor_key(c)
    n = 131                                                  # 64 (Fq2) / 40 (Fq3) points per block: three or four blocks, ragged tail
    base = [G.mul(G.gen, rnd.randrange(1, G.r)) for _ in range(4)]
    pts = [base[i % 4] for i in range(n)]
    pts[77] = None
    first = (1 << 24) + 7
    ks = [key.alpha * pow(key.tau, first + j, c.Fr.p) % c.Fr.p for j in range(n)]
    # the oracle on the points that differ in (point, scalar); every output is checked against the other body bit for bit
    d_in = dev_bytes(ser.points_to_bytes(G, pts, False))
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SSO_COOP_G2", mode)
        d_out = torch.zeros(n * ser.point_size(G, True), dtype=torch.uint8, device="cuda")
        sso.batch_exp(name, 1, d_in, n, first, key.tau, key.alpha, d_out)
        outs[mode] = host_bytes(d_out)
        with pytest.raises(sso.SsoError) as e:
            sso.batch_exp(name, 1, d_in, n, first, key.tau, key.alpha, d_out, check=sso.CHECK_NONZERO)
        assert e.value.code == -3 and "infinity" in e.value.message and "77" in e.value.message
    assert outs["0"] == outs["1"]
    sz = ser.point_size(G, True)
    for j in (0, 1, 63, 64, 76, 77, 78, 119, 120, 130):
        assert outs["1"][j * sz:(j + 1) * sz] == ser.point_to_bytes(G, G.mul(pts[j], ks[j]) if pts[j] is not None else None, True), j
    # CHECK_FULL (on-curve test of uncompressed input) accepts, and rejects a point off the curve, in both bodies
    good = [p for p in pts if p is not None][:70]
    bad = list(good)
    bad[41] = (bad[41][0], G.F.add(bad[41][1], G.F.one))
    for mode in ("0", "1"):
        monkeypatch.setenv("SSO_COOP_G2", mode)
        d_out = torch.zeros(70 * sz, dtype=torch.uint8, device="cuda")
        sso.batch_exp(name, 1, dev_bytes(ser.points_to_bytes(G, good, False)), 70, 0, key.tau, None, d_out, check=sso.CHECK_FULL)
        with pytest.raises(sso.SsoError) as e:
            sso.batch_exp(name, 1, dev_bytes(ser.points_to_bytes(G, bad, False)), 70, 0, key.tau, None, d_out, check=sso.CHECK_FULL)
        assert "41" in e.value.message
    # a whole chunk contribution (fused G1 + G2 launch) through both bodies
    from oracle.params import Phase1Params
    o = Phase1Params.new_chunk(name, 1, 8, 4, 8)
    p = sso.Phase1Parameters.new_chunk(name, 1, 8, 4, 8)
    ch = synth.synthetic_challenge(o)
    resp = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SSO_COOP_G2", mode)
        r = bytearray(p.contribution_size)
        sso.contribute_buf(p, ch, r, key.tau, key.alpha, key.beta, pubkey=bytes(o.public_key_size))
        resp[mode] = bytes(r)
    assert resp["0"] == resp["1"]
    from oracle import phase1
    assert resp["1"] == phase1.contribute_with_key(o, ch, key, bytes(o.public_key_size))


@pytest.mark.parametrize("name", CURVE_NAMES)
@pytest.mark.parametrize("gi", [0, 1])
def test_batch_exp_matches_oracle(name, gi):
    c = get_curve(name)
    G = (c.g1, c.g2)[gi]
    rnd = random.Random(3 * c.cid + gi)
    n = 5
    key = synth.contributor_key(c)
    pts = [G.mul(G.gen, rnd.randrange(1, G.r)) for _ in range(n - 1)] + [None]
    first = (1 << 20) + 12345
    want = ser.points_to_bytes(G, [G.mul(P, key.alpha * pow(key.tau, first + j, c.Fr.p) % c.Fr.p) for j, P in enumerate(pts)], True)
    d_in = dev_bytes(ser.points_to_bytes(G, pts, False))
    d_out = torch.zeros(len(want), dtype=torch.uint8, device="cuda")
    sso.batch_exp(name, gi, d_in, n, first, key.tau, key.alpha, d_out)
    assert host_bytes(d_out) == want
    # CHECK_NONZERO reports the point at infinity as an input error
    with pytest.raises(sso.SsoError) as e:
        sso.batch_exp(name, gi, d_in, n, first, key.tau, key.alpha, d_out, check=sso.CHECK_NONZERO)
    assert e.value.code == -3 and "infinity" in e.value.message
    # phase-2 style shared scalar, uncompressed output
    want2 = ser.points_to_bytes(G, [G.mul(P, key.beta) for P in pts], False)
    d_out2 = torch.zeros(len(want2), dtype=torch.uint8, device="cuda")
    sso.batch_mul(name, gi, d_in, n, key.beta, d_out2, out_compressed=False)
    assert host_bytes(d_out2) == want2
    # decompress what we produced, with curve + subgroup checks
    d_back = torch.zeros((n - 1) * ser.point_size(G, False), dtype=torch.uint8, device="cuda")
    sso.reencode(name, gi, d_out[:(n - 1) * ser.point_size(G, True)].contiguous(), n - 1, d_back)
    pts_back = ser.points_from_bytes(G, host_bytes(d_back), False)
    assert ser.points_to_bytes(G, pts_back, True) == want[:(n - 1) * ser.point_size(G, True)]


@pytest.mark.parametrize("name", CURVE_NAMES)
@pytest.mark.parametrize("k", [0, 3])
def test_contribute_matches_golden_chunks(name, k, golden_dir):
    p = sso.Phase1Parameters.new_chunk(name, k, 4, 3, 4)
    ch = open(os.path.join(golden_dir, "p1_%s_c%d.challenge.bin" % (name, k)), "rb").read()
    want = open(os.path.join(golden_dir, "p1_%s_c%d.response.bin" % (name, k)), "rb").read()
    key = synth.contributor_key(get_curve(name))
    resp = bytearray(p.contribution_size)
    sso.contribute_buf(p, ch, resp, key.tau, key.alpha, key.beta, pubkey=bytes(p.sizes()["public_key_size"]),
                       check=sso.CHECK_NONZERO)
    assert bytes(resp) == want                          # includes the Blake2b hash-chain link in [0, 64)


def test_contribute_many_chunks_in_flight_matches_golden(golden_dir):
    """sso_p1_contribute_many_buf (several chunks in flight on host worker threads) = the single-chunk call per chunk."""
    name = "bls12_377"
    key = synth.contributor_key(get_curve(name))
    order = [0, 3, 3, 0, 3, 0, 0]
    params = [sso.Phase1Parameters.new_chunk(name, k, 4, 3, 4) for k in order]
    chs = [open(os.path.join(golden_dir, "p1_%s_c%d.challenge.bin" % (name, k)), "rb").read() for k in order]
    wants = [open(os.path.join(golden_dir, "p1_%s_c%d.response.bin" % (name, k)), "rb").read() for k in order]
    resps = [bytearray(p.contribution_size) for p in params]
    sso.contribute_many_buf(params, chs, resps, key.tau, key.alpha, key.beta, pubkey=bytes(params[0].sizes()["public_key_size"]),
                            check=sso.CHECK_NONZERO, host_threads=3)
    assert [bytes(r) for r in resps] == wants
    # an invalid chunk in the batch surfaces as that chunk's error
    bad = chs[2][:-1]                                    # wrong length
    with pytest.raises(sso.SsoError) as e:
        sso.contribute_many_buf(params, chs[:2] + [bytes(bad)] + chs[3:], resps, key.tau, key.alpha, key.beta,
                                pubkey=bytes(params[0].sizes()["public_key_size"]), check=sso.CHECK_NONZERO, host_threads=2)
    assert "chunk 2 of the batch" in str(e.value)
    sso.contribute_many_buf([], [], [], 1, 2, 3)


def test_reencode_rejects_bad_points():
    c = get_curve("bls12_377")
    G = c.g1
    from oracle.curves import _some_point
    rogue = _some_point(G, 11)
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    with pytest.raises(sso.SsoError) as e:
        sso.reencode("bls12_377", 0, dev_bytes(ser.point_to_bytes(G, rogue, True)), 1, d_out)
    assert e.value.code == -4 and "subgroup" in e.value.message
    sso.reencode("bls12_377", 0, dev_bytes(ser.point_to_bytes(G, rogue, True)), 1, d_out, subgroup_check=False)
    assert host_bytes(d_out) == ser.point_to_bytes(G, rogue, False)
    bad = bytearray(ser.point_to_bytes(G, G.gen, True)); bad[-1] |= 0xC0
    with pytest.raises(sso.SsoError) as e:
        sso.reencode("bls12_377", 0, dev_bytes(bad), 1, d_out)
    assert e.value.code == -3 and "flags" in e.value.message
    with pytest.raises(sso.SsoError) as e:
        sso.reencode("bls12_377", 0, dev_bytes(c.Fq.p.to_bytes(48, "little")), 1, d_out)
    assert "canonical" in e.value.message


def test_wrong_sizes_are_argument_errors():
    p = sso.Phase1Parameters.new_chunk("bls12_377", 0, 4, 3, 4)
    with pytest.raises(sso.SsoError) as e:
        sso.contribute_buf(p, bytes(p.accumulator_size - 1), bytearray(p.contribution_size), 1, 2, 3)
    assert e.value.code == -1 and "accumulator_size" in e.value.message


def _decompress_response(p, d_resp):
    """response vectors (compressed, device) -> new challenge image (uncompressed, device)"""
    o = Phase1Params.new_chunk(p.curve, p.chunk_index, p.chunk_size, p.power, p.batch_size)
    oc, ou = o.offsets(True), o.offsets(False)
    d_new = torch.zeros(o.accumulator_size, dtype=torch.uint8, device="cuda")
    counts = (o.g1_count, o.other_count, o.other_count, o.other_count, 1)
    groups = (0, 1, 0, 0, 1)
    for i in range(5):
        if counts[i]:
            sso.reencode(p.curve, groups[i], d_resp[oc[i]:oc[i + 1]], counts[i], d_new[ou[i]:ou[i + 1]])
    return d_new


@pytest.mark.parametrize("name,clog", [("bls12_377", 12), ("bw6_761", 8), ("mnt4_753", 8), ("mnt6_753", 7)])
def test_composition_property_at_size(name, clog):
    """contribute(contribute(X, k1), k2) == contribute(X, k1 * k2) byte for byte, on a chunk far
    larger than the oracle can handle; a handful of elements are also checked against the oracle."""
    c = get_curve(name)
    cs = 1 << clog
    p = sso.Phase1Parameters.new_chunk(name, 1, cs, clog + 2, cs)
    o = Phase1Params.new_chunk(name, 1, cs, clog + 2, cs)
    # generator accumulator, built with the oracle's serializer
    d_gen = dev_bytes(phase1.new_challenge(o))
    k1 = phase1.PrivateKey(*synth.scalars_from_seed(c, synth.SEED_PREV))
    k2 = synth.contributor_key(c)
    r = c.Fr.p
    k12 = phase1.PrivateKey(k1.tau * k2.tau % r, k1.alpha * k2.alpha % r, k1.beta * k2.beta % r)
    d_r1 = torch.zeros(o.contribution_size, dtype=torch.uint8, device="cuda")
    sso.contribute_dev(p, d_gen, d_r1, k1.tau, k1.alpha, k1.beta)
    d_c1 = _decompress_response(p, d_r1)
    d_r2 = torch.zeros(o.contribution_size, dtype=torch.uint8, device="cuda")
    sso.contribute_dev(p, d_c1, d_r2, k2.tau, k2.alpha, k2.beta)
    d_r12 = torch.zeros(o.contribution_size, dtype=torch.uint8, device="cuda")
    sso.contribute_dev(p, d_gen, d_r12, k12.tau, k12.alpha, k12.beta)
    end = o.offsets(True)[5]
    assert torch.equal(d_r2[64:end], d_r12[64:end])
    # spot-check against the oracle: first / last element of every vector of the first contribution
    resp = host_bytes(d_r1)
    oc = o.offsets(True)
    g1c, g2c = o.sizes(True)
    for vec, G, size, coeff, cnt in ((0, c.g1, g1c, 1, o.g1_count), (1, c.g2, g2c, 1, o.other_count),
                                     (2, c.g1, g1c, k1.alpha, o.other_count), (3, c.g1, g1c, k1.beta, o.other_count)):
        for j in (0, cnt - 1):
            want = G.mul(G.gen, coeff * pow(k1.tau, o.start + j, r) % r)
            assert resp[oc[vec] + j * size: oc[vec] + (j + 1) * size] == ser.point_to_bytes(G, want, True)
    assert resp[oc[4]:oc[5]] == ser.point_to_bytes(c.g2, c.g2.mul(c.g2.gen, k1.beta), True)


def test_imad_probe_runs():
    out = ctypes.c_double(0)
    for variant in (0, 1, 2):
        sso._lib.call("sso_imad_peak", 0, variant, ctypes.byref(out))
        assert out.value > 1e11


@pytest.mark.parametrize("name,gi,n", [("bls12_377", 0, 300), ("bls12_377", 1, 70), ("bw6_761", 0, 40), ("mnt4_753", 1, 20),
                                        ("mnt6_753", 1, 12)])
def test_power_and_merge_pairs_match_oracle(name, gi, n):
    """Pippenger MSM on the GPU (CUB-sorted buckets) against the oracle's plain sum r_i P_i with the same
    ChaCha20-derived scalars; inputs are decompressed and checked on the way in."""
    c = get_curve(name)
    G = (c.g1, c.g2)[gi]
    rnd = random.Random(n)
    base = [G.mul(G.gen, rnd.randrange(1, G.r)) for _ in range(6)]
    pts = [base[i % 6] if i % 7 else G.add(base[i % 6], base[(i + 1) % 6]) for i in range(n)]
    seed = bytes(range(7, 39))
    rs = phase1.rlc_scalars(c, seed, n)
    a, b = phase1.power_pairs_with(G, pts, rs[:n - 1])
    got = sso.power_pairs(name, gi, dev_bytes(ser.points_to_bytes(G, pts, True)), n, in_compressed=True, check=sso.CHECK_FULL,
                          subgroup_check=True, seed32=seed)
    assert got == ser.point_to_bytes(G, a, False) + ser.point_to_bytes(G, b, False)
    pts2 = [G.mul(P, 3) for P in pts]
    a2, b2 = phase1.merge_pairs_with(G, pts, pts2, rs)
    got2 = sso.merge_pairs(name, gi, dev_bytes(ser.points_to_bytes(G, pts, False)), dev_bytes(ser.points_to_bytes(G, pts2, False)), n,
                           seed32=seed)
    assert got2 == ser.point_to_bytes(G, a2, False) + ser.point_to_bytes(G, b2, False)
    assert G.eq(G.mul(a2, 3), b2)                        # the pair has the ratio the vectors have


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_same_ratio_verdicts(name):
    c = get_curve(name)
    g1, g2 = c.g1, c.g2
    P, Q = g1.mul(g1.gen, 0x1234567), g2.mul(g2.gen, 0x7654321)
    x = 0xABCDEF0123456789

    def u(G, pt):
        return ser.point_to_bytes(G, pt, False)

    checks = [(u(g1, P), u(g1, g1.mul(P, x)), u(g2, Q), u(g2, g2.mul(Q, x))),
              (u(g1, P), u(g1, g1.mul(P, x)), u(g2, Q), u(g2, g2.mul(Q, x + 1))),
              (u(g1, g1.gen), u(g1, g1.gen), u(g2, g2.gen), u(g2, g2.gen))]
    assert sso.same_ratio(name, checks) == [True, False, True]
    bad = bytearray(checks[0][0]); bad[3] ^= 1                      # not on the curve any more
    with pytest.raises(sso.SsoError) as e:
        sso.same_ratio(name, [(bytes(bad),) + checks[0][1:]])
    assert e.value.code == -3


@pytest.mark.parametrize("name,power,clog,k", [("bls12_377", 20, 16, 1),      # BASELINE config 2: one full chunk
                                               ("bw6_761", 12, 10, 0),        # BASELINE config 1: first chunk
                                               ("bw6_761", 12, 10, 7)])       # BASELINE config 1: G1-only tail chunk
def test_named_configs_at_full_size(name, power, clog, k):
    """The configurations BASELINE.json names, at their real sizes: composition of two contributions equals the
    contribution with the product key (byte for byte), the verifier accepts the result, and sampled elements match
    the oracle's scalar multiplication."""
    c = get_curve(name)
    cs = 1 << clog
    p = sso.Phase1Parameters.new_chunk(name, k, cs, power, cs)
    o = Phase1Params.new_chunk(name, k, cs, power, cs)
    d_gen = torch.empty(o.accumulator_size, dtype=torch.uint8, device="cuda")
    sso.new_challenge_dev(p, d_gen)
    r = c.Fr.p
    k1 = phase1.PrivateKey(*synth.scalars_from_seed(c, synth.SEED_PREV))
    k2 = synth.contributor_key(c)
    k12 = phase1.PrivateKey(k1.tau * k2.tau % r, k1.alpha * k2.alpha % r, k1.beta * k2.beta % r)
    d_r1 = torch.zeros(o.contribution_size, dtype=torch.uint8, device="cuda")
    sso.contribute_dev(p, d_gen, d_r1, k1.tau, k1.alpha, k1.beta)
    d_c1 = _decompress_response(p, d_r1)
    d_r2 = torch.zeros(o.contribution_size, dtype=torch.uint8, device="cuda")
    sso.contribute_dev(p, d_c1, d_r2, k2.tau, k2.alpha, k2.beta)
    d_r12 = torch.zeros(o.contribution_size, dtype=torch.uint8, device="cuda")
    sso.contribute_dev(p, d_gen, d_r12, k12.tau, k12.alpha, k12.beta)
    end = o.offsets(True)[5]
    assert torch.equal(d_r2[64:end], d_r12[64:end])
    resp = host_bytes(d_r1)
    oc = o.offsets(True)
    g1c, g2c = o.sizes(True)
    rnd = random.Random(k)
    for vec, G, size, coeff, cnt in ((0, c.g1, g1c, 1, o.g1_count), (1, c.g2, g2c, 1, o.other_count),
                                     (2, c.g1, g1c, k1.alpha, o.other_count), (3, c.g1, g1c, k1.beta, o.other_count)):
        if cnt == 0:
            continue
        for j in {0, cnt - 1, rnd.randrange(cnt)}:
            want = G.mul(G.gen, coeff * pow(k1.tau, o.start + j, r) % r)
            assert resp[oc[vec] + j * size: oc[vec] + (j + 1) * size] == ser.point_to_bytes(G, want, True)
    # a full seeded contribution of the same chunk is accepted by the verifier (hash chain, PoK, ratio checks)
    ch = host_bytes(d_c1)
    h_ch = bytearray(ch)
    h_ch[:64] = phase1.blank_hash()
    resp2 = bytearray(o.contribution_size)
    sso.contribute_seeded_buf(p, bytes(h_ch), resp2, synth.SEED_CONTRIB)
    new_ch = bytearray(o.accumulator_size)
    sso.verify_chunk_buf(p, bytes(h_ch), bytes(resp2), new_ch)
    assert bytes(new_ch[:64]) == phase1.calculate_hash(bytes(resp2))
