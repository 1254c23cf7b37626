"""Phase 2 on the Groth16 parameter container through the C ABI (rows a10, a11): phase2_cli::contribute / verify restated
by oracle/phase2.py (keys and pairings in Python big integers) against sso_p2_contribute_* / sso_p2_verify_*: response and
new-challenge bytes, a chained second contribution, rejects, the file-level calls with the reference's argument order."""
import hashlib
import random

import pytest

import snark_setup_operator_b200 as sso
from oracle import phase2 as o2, serialize as ser
from oracle.chacha import ChaChaRng
from oracle.curves import get_curve
from snark_setup_operator_b200 import phase2 as p2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SEED1, SEED2 = bytes(range(32)), bytes(range(100, 132))


def _params(c, rnd, nh=7, nl=5):
    r = c.Fr.p
    p1 = lambda: c.g1.mul(c.g1.gen, rnd.randrange(1, r))
    q2 = lambda: c.g2.mul(c.g2.gen, rnd.randrange(1, r))
    a_query = [p1() for _ in range(5)]
    a_query[2] = None                                        # queries may hold the point at infinity
    return o2.MPCParameters(p1(), q2(), q2(), c.g2.gen, [p1() for _ in range(3)], p1(), c.g1.gen, a_query, [p1() for _ in range(4)],
                            [q2() for _ in range(4)], [p1() for _ in range(nh)], [p1() for _ in range(nl)],
                            hashlib.blake2b(b"constraint system", digest_size=64).digest())


@pytest.mark.parametrize("name", ["bls12_377", "mnt4_753", "mnt6_753", "bw6_761"])
def test_phase2_contribute_and_verify(name):
    c = get_curve(name)
    m = _params(c, random.Random(7))
    ch = m.to_bytes(c, False)
    resp = p2.contribute_buf(name, ch, SEED1)
    assert resp == o2.contribute(c, ch, ChaChaRng(SEED1))
    new = p2.verify_buf(name, ch, resp, rlc_seed32=bytes(32))
    assert new == o2.verify(c, ch, resp)
    # a second contribution chains on the new challenge (transcript over the earlier public key)
    resp2 = p2.contribute_buf(name, new, SEED2)
    assert resp2 == o2.contribute(c, new, ChaChaRng(SEED2))
    new2 = p2.verify_buf(name, new, resp2)
    assert o2.MPCParameters.from_bytes(c, new2, False).contributions[0] == o2.MPCParameters.from_bytes(c, resp, True).contributions[0]
    # delta_g1 / delta_g2 moved by delta1 * delta2, h / l by its inverse
    from oracle.chacha import fp_rand
    d1, d2 = fp_rand(c.Fr, ChaChaRng(SEED1)), fp_rand(c.Fr, ChaChaRng(SEED2))
    fin = o2.MPCParameters.from_bytes(c, new2, False)
    assert c.g1.eq(fin.delta_g1, c.g1.mul(c.g1.gen, d1 * d2 % c.Fr.p))
    inv = pow(d1 * d2, -1, c.Fr.p)
    assert all(c.g1.eq(P, c.g1.mul(Q, inv)) for P, Q in zip(fin.h_query, m.h_query))


def test_phase2_rejects():
    name = "bls12_377"
    c = get_curve(name)
    m = _params(c, random.Random(8))
    ch = m.to_bytes(c, False)
    resp = p2.contribute_buf(name, ch, SEED1)
    good = o2.MPCParameters.from_bytes(c, resp, True)

    def rejected(mod, match=None):
        bad = o2.MPCParameters.from_bytes(c, resp, True)
        mod(bad)
        with pytest.raises(sso.SsoError) as e:
            p2.verify_buf(name, ch, bad.to_bytes(c, True))
        assert e.value.code == -4, e.value.message
        if match:
            assert match in e.value.message, e.value.message
        with pytest.raises(ValueError):
            o2.verify(c, ch, bad.to_bytes(c, True))

    def tamper_h(b): b.h_query[3] = c.g1.mul(c.g1.gen, 99)
    def tamper_l(b): b.l_query[0] = c.g1.mul(b.l_query[0], 2)
    def tamper_a(b): b.a_query[1] = c.g1.mul(c.g1.gen, 5)
    def tamper_delta(b): b.delta_g2 = c.g2.mul(b.delta_g2, 3)
    def tamper_cs(b): b.cs_hash = bytes(64)
    def tamper_pk(b): b.contributions[-1] = b.contributions[-1][:-1] + bytes([b.contributions[-1][-1] ^ 1])
    rejected(tamper_h, "h_query")
    rejected(tamper_l, "l_query")
    rejected(tamper_a, "changed")
    rejected(tamper_delta, "delta")
    rejected(tamper_cs, "cs_hash")
    rejected(tamper_pk, "transcript")
    assert p2.verify_buf(name, ch, good.to_bytes(c, True))


def test_phase2_file_calls(tmp_path):
    name = "mnt4_753"
    c = get_curve(name)
    ch = _params(c, random.Random(9)).to_bytes(c, False)
    f = {k: str(tmp_path / k) for k in ("challenge", "challenge.hash", "response", "response.hash", "c.vhash", "r.vhash", "new_challenge", "new_challenge.hash")}
    open(f["challenge"], "wb").write(ch)
    p2.contribute(name, f["challenge"], f["challenge.hash"], f["response"], f["response.hash"], sso.CHECK_NO, 0, SEED1)
    resp = open(f["response"], "rb").read()
    assert resp == o2.contribute(c, ch, ChaChaRng(SEED1))
    b2 = lambda d: hashlib.blake2b(d, digest_size=64).digest()
    assert open(f["challenge.hash"], "rb").read() == b2(ch) and open(f["response.hash"], "rb").read() == b2(resp)
    p2.verify(name, f["challenge"], f["c.vhash"], sso.CHECK_NO, f["response"], f["r.vhash"], sso.CHECK_NO, f["new_challenge"], f["new_challenge.hash"], 0, False)
    new = open(f["new_challenge"], "rb").read()
    assert new == o2.MPCParameters.from_bytes(c, resp, True).to_bytes(c, False)
    assert open(f["new_challenge.hash"], "rb").read() == b2(new) and open(f["r.vhash"], "rb").read() == b2(resp)
    with pytest.raises(sso.SsoError) as e:                      # outputs must not exist
        p2.contribute(name, f["challenge"], str(tmp_path / "x.hash"), f["response"], str(tmp_path / "y.hash"), sso.CHECK_NO, 0, SEED1)
    assert e.value.code == -5


def test_config4_query_of_2_19_points_spot_checked():
    """BASELINE config 4 at a Nimiq circuit size (2^19 G1 points, reference e2e/nimiq_e2e.sh:61-71): the delta^-1 scaling of the
    whole vector on the GPU, sampled elements recomputed one by one by the C++ oracle leg (plain double-and-add), and the
    same-ratio check of the whole vector against (delta_g2_after, delta_g2_before)."""
    from oracle import cport, synth
    name = "mnt4_753"
    c = get_curve(name)
    n = 1 << 19
    es = sso.phase1.curve_sizes(name)
    usz = es["g1_u"]
    s = synth.scalars_from_seed(c, synth.SEED_PREV)[0]
    delta = synth.scalars_from_seed(c, synth.SEED_CONTRIB, 4)[3]
    dinv = pow(delta, -1, c.Fr.p)
    gen = ser.point_to_bytes(c.g1, c.g1.gen, False)
    d_gen = torch.frombuffer(bytearray(gen * n), dtype=torch.uint8).cuda()
    d_c = torch.empty(n * es["g1_c"], dtype=torch.uint8, device="cuda")
    sso.batch_exp(name, 0, d_gen, n, 1, s, None, d_c)                       # h_i = s^(1+i) G
    d_u = torch.empty(n * usz, dtype=torch.uint8, device="cuda")
    sso.reencode(name, 0, d_c, n, d_u, check=sso.CHECK_NO, subgroup_check=False)
    before = d_u.cpu().numpy().tobytes()
    after = p2.scale_queries(name, before, n, dinv)
    rnd = random.Random(3)
    for j in [0, 1, n - 1] + [rnd.randrange(n) for _ in range(5)]:
        want = cport.batch_exp(c, 0, before[j * usz:(j + 1) * usz], 1, 0, 1, dinv, mode=1, out_compressed=False, threads=1)
        assert after[j * usz:(j + 1) * usz] == want, j
    d2b = c.g2.mul(c.g2.gen, 31337)
    d2a = c.g2.mul(d2b, delta)
    p2.verify_queries(name, before, after, n, ser.point_to_bytes(c.g2, d2b, False), ser.point_to_bytes(c.g2, d2a, False))
    bad = bytearray(after)
    bad[12345 * usz:12346 * usz] = before[12345 * usz:12346 * usz]
    with pytest.raises(sso.SsoError) as e:
        p2.verify_queries(name, before, bytes(bad), n, ser.point_to_bytes(c.g2, d2b, False), ser.point_to_bytes(c.g2, d2a, False))
    assert e.value.code == -4
