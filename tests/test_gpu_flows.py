"""The file-level ceremony calls on the GPU through the C ABI: key generation from the seeded RNG, the full
phase1_cli::contribute (hash chain + public key), chunk verification with real pairings (accept / reject
verdicts against the oracle), and the file entry points with the reference's argument order."""
import os

import pytest

import snark_setup_operator_b200 as sso
from oracle import phase1, serialize as ser, synth
from oracle.chacha import ChaChaRng
from oracle.curves import CURVE_NAMES, get_curve
from oracle.params import Phase1Params

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_keygen_matches_oracle(name):
    c = get_curve(name)
    digest = phase1.calculate_hash(b"some challenge")
    pub, key = phase1.key_generation(c, ChaChaRng(synth.SEED_CONTRIB), digest)
    scalars, pk = sso.keygen(name, synth.SEED_CONTRIB, digest)
    assert scalars == (key.tau, key.alpha, key.beta)
    assert pk == pub.to_bytes(c)


def _contribution(name, k, cs=4, power=3):
    o = Phase1Params.new_chunk(name, k, cs, power, cs)
    p = sso.Phase1Parameters.new_chunk(name, k, cs, power, cs)
    ch = synth.synthetic_challenge(o)
    resp = bytearray(o.contribution_size)
    sso.contribute_seeded_buf(p, ch, resp, synth.SEED_CONTRIB)
    return o, p, ch, bytes(resp)


def test_full_contribute_matches_oracle():
    o, p, ch, resp = _contribution("bls12_377", 1)
    want, ch_hash, resp_hash = phase1.contribute(o, ch, ChaChaRng(synth.SEED_CONTRIB))
    assert resp == want


@pytest.mark.parametrize("name,k", [("bls12_377", 0), ("bls12_377", 1), ("bls12_377", 3), ("bw6_761", 1), ("mnt4_753", 0),
                                    ("mnt6_753", 1)])
def test_verify_accepts_honest_contribution(name, k):
    o, p, ch, resp = _contribution(name, k)
    new_ch = bytearray(o.accumulator_size)
    sso.verify_chunk_buf(p, ch, resp, new_ch, rlc_seed32=bytes(range(32)))
    assert bytes(new_ch) == phase1.decompress_response(o, resp)
    if name == "bls12_377" and k == 1:
        assert phase1.verify_chunk(o, ch, resp, rlc_seed32=bytes(range(32))) == bytes(new_ch)


def test_full_contribute_many_chunks_in_flight():
    """sso_p1_contribute_seeded_many_buf = phase1_cli::contribute per chunk (same seed-derived key, per-chunk proofs of knowledge)"""
    ks = [0, 1, 3, 1]
    os_ = [Phase1Params.new_chunk("bls12_377", k, 4, 3, 4) for k in ks]
    ps = [sso.Phase1Parameters.new_chunk("bls12_377", k, 4, 3, 4) for k in ks]
    chs = [synth.synthetic_challenge(o) for o in os_]
    resps = [bytearray(o.contribution_size) for o in os_]
    sso.contribute_seeded_many_buf(ps, chs, resps, synth.SEED_CONTRIB, host_threads=4)
    for o, ch, r in zip(os_, chs, resps):
        want, _, _ = phase1.contribute(o, ch, ChaChaRng(synth.SEED_CONTRIB))
        assert bytes(r) == want


def test_verify_many_chunks_in_flight():
    """sso_p1_verify_chunk_many_buf: the chunk loop of verify_transcript as a work queue — same new challenges as the
    single-chunk call, and a bad chunk in the batch is reported with its index."""
    items = [_contribution("bls12_377", k) for k in (0, 1, 3, 1, 0)]
    params = [it[1] for it in items]
    chs = [it[2] for it in items]
    resps = [it[3] for it in items]
    news = [bytearray(it[0].accumulator_size) for it in items]
    sso.verify_chunk_many_buf(params, chs, resps, news, rlc_seed32=bytes(range(32)), host_threads=3)
    for it, new in zip(items, news):
        assert bytes(new) == phase1.decompress_response(it[0], it[3])
    sso.verify_chunk_many_buf(params, chs, resps, news, rlc_seed32=bytes(range(32)), host_threads=0, device=-1)   # all devices
    bad = bytearray(resps[3])
    bad[-1] ^= 1                                                  # corrupt the public key of chunk 3
    with pytest.raises(sso.SsoError) as e:
        sso.verify_chunk_many_buf(params, chs, resps[:3] + [bytes(bad)] + resps[4:], news, rlc_seed32=bytes(range(32)), host_threads=2)
    assert e.value.code in (-3, -4) and "chunk 3 of the batch" in str(e.value)


def test_verify_rejects_bad_contributions():
    o, p, ch, resp = _contribution("bls12_377", 0)
    c = o.curve
    new_ch = bytearray(o.accumulator_size)
    oc = o.offsets(True)

    def rejected(bad, match, **kw):
        with pytest.raises(sso.SsoError) as e:
            sso.verify_chunk_buf(p, ch, bytes(bad), new_ch, **kw)
        assert e.value.code == -4 and match in e.value.message, e.value.message

    bad = bytearray(resp); bad[0] ^= 1
    rejected(bad, "hash chain")
    # a valid subgroup point with the wrong scalar in the middle of tau_g1: only the RLC power check sees it
    bad = bytearray(resp); bad[oc[0] + 2 * 48: oc[0] + 3 * 48] = ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 12345), True)
    rejected(bad, "power ratio: tau_g1")
    with pytest.raises(phase1.VerificationError, match="power ratio: tau_g1"):
        phase1.verify_chunk(o, ch, bytes(bad))
    sso.verify_chunk_buf(p, ch, bytes(bad), new_ch, ratio_check=False)      # --skip-ratio-check lets it through
    # alpha_g1[0] replaced: the before/after check against the alpha proof fails
    bad = bytearray(resp); bad[oc[2]: oc[2] + 48] = ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 777), True)
    rejected(bad, "alpha")
    # a point outside the prime-order subgroup
    from oracle.curves import _some_point
    bad = bytearray(resp); bad[oc[3] + 48: oc[3] + 96] = ser.point_to_bytes(c.g1, _some_point(c.g1, 11), True)
    rejected(bad, "subgroup")
    # tampered public key: proof of knowledge fails
    bad = bytearray(resp); pk0 = oc[5]
    bad[pk0 + 96: pk0 + 192] = ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 99), False)
    rejected(bad, "proof of knowledge: tau")
    # point at infinity in the output
    bad = bytearray(resp); bad[oc[0] + 48: oc[0] + 96] = ser.point_to_bytes(c.g1, None, True)
    rejected(bad, "infinity")
    # tau_g1[0] moved away from the generator
    bad = bytearray(resp); bad[oc[0]: oc[0] + 48] = ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 2), True)
    rejected(bad, "generator", ratio_check=False)


def test_file_entry_points(tmp_path):
    """contribute -> transform_pok_and_correctness through files, reference argument order; hash files hold the
    raw 64 bytes (reference src/utils.rs:264-276); existing outputs are an error (create_new)."""
    name = "bls12_377"
    o = Phase1Params.new_chunk(name, 2, 4, 3, 4)
    p = sso.Phase1Parameters.new_chunk(name, 2, 4, 3, 4)
    f = {k: str(tmp_path / k) for k in ("challenge", "challenge.hash", "response", "response.hash", "new_challenge",
                                         "new_challenge.hash", "challenge.verified.hash", "response.verified.hash")}
    ch = synth.synthetic_challenge(o)
    open(f["challenge"], "wb").write(ch)
    sso.contribute(f["challenge"], f["challenge.hash"], f["response"], f["response.hash"], sso.CHECK_NONZERO, 0, p, synth.SEED_CONTRIB)
    resp = open(f["response"], "rb").read()
    want, ch_hash, resp_hash = phase1.contribute(o, ch, ChaChaRng(synth.SEED_CONTRIB))
    assert resp == want
    assert open(f["challenge.hash"], "rb").read() == ch_hash and open(f["response.hash"], "rb").read() == resp_hash
    with pytest.raises(sso.SsoError) as e:                      # outputs already exist
        sso.contribute(f["challenge"], f["challenge.hash"], f["response"], f["response.hash"], sso.CHECK_NONZERO, 0, p, synth.SEED_CONTRIB)
    assert e.value.code == -5
    sso.transform_pok_and_correctness(f["challenge"], f["challenge.verified.hash"], sso.CHECK_NO, f["response"], f["response.verified.hash"],
                                      sso.CHECK_FULL, f["new_challenge"], f["new_challenge.hash"], 0, True, p)
    new_ch = open(f["new_challenge"], "rb").read()
    assert new_ch == phase1.decompress_response(o, resp)
    assert open(f["new_challenge.hash"], "rb").read() == phase1.calculate_hash(new_ch)
    assert open(f["response.verified.hash"], "rb").read() == resp_hash
    # wrong-size input file: the reference asserts file lengths
    open(f["challenge"], "ab").write(b"\0")
    with pytest.raises(sso.SsoError) as e:
        sso.contribute(f["challenge"], str(tmp_path / "x1"), str(tmp_path / "x2"), str(tmp_path / "x3"), 0, 0, p, synth.SEED_CONTRIB)
    assert e.value.code == -1 and "size" in e.value.message


def test_failed_file_call_leaves_nothing_behind(tmp_path):
    """no half-written outputs — a call that fails after it created its outputs (here: an off-subgroup point in
    the challenge under CheckForCorrectness::Full, a response that fails verification) removes them, also the hash files."""
    from oracle.curves import _some_point
    name = "bls12_377"
    o = Phase1Params.new_chunk(name, 1, 4, 3, 4)
    p = sso.Phase1Parameters.new_chunk(name, 1, 4, 3, 4)
    c = o.curve
    ch = bytearray(synth.synthetic_challenge(o))
    ou = o.offsets(False)
    ch[ou[0] + 96: ou[0] + 192] = ser.point_to_bytes(c.g1, _some_point(c.g1, 11), False)
    f = {k: str(tmp_path / k) for k in ("challenge", "challenge.hash", "response", "response.hash", "new", "new.hash", "c.vhash", "r.vhash")}
    open(f["challenge"], "wb").write(ch)
    with pytest.raises(sso.SsoError):
        sso.contribute(f["challenge"], f["challenge.hash"], f["response"], f["response.hash"], sso.CHECK_FULL, 0, p, synth.SEED_CONTRIB)
    assert sorted(os.listdir(tmp_path)) == ["challenge"]
    # the same for a Full-mode (streamed, memory-mapped) call and for a rejected verification
    pf = sso.Phase1Parameters.new_full(name, 3, 4)
    of = Phase1Params.new_full(name, 3, 4)
    chf = bytearray(phase1.new_challenge(of))
    ouf = of.offsets(False)
    chf[ouf[2] + 96: ouf[2] + 192] = ser.point_to_bytes(c.g1, _some_point(c.g1, 11), False)
    open(f["challenge"], "wb").write(chf)
    with pytest.raises(sso.SsoError):
        sso.contribute(f["challenge"], f["challenge.hash"], f["response"], f["response.hash"], sso.CHECK_FULL, 0, pf, synth.SEED_CONTRIB)
    assert sorted(os.listdir(tmp_path)) == ["challenge"]
    open(f["challenge"], "wb").write(phase1.new_challenge(of))
    sso.contribute(f["challenge"], f["challenge.hash"], f["response"], f["response.hash"], sso.CHECK_NONZERO, 0, pf, synth.SEED_CONTRIB)
    resp = bytearray(open(f["response"], "rb").read())
    resp[-1] ^= 1                                               # public key tampered: verification fails at the end of the call
    open(f["response"], "wb").write(resp)
    with pytest.raises(sso.SsoError) as e:
        sso.transform_pok_and_correctness(f["challenge"], f["c.vhash"], sso.CHECK_NO, f["response"], f["r.vhash"], sso.CHECK_NONZERO, f["new"],
                                          f["new.hash"], 0, True, pf)
    assert e.value.code in (-3, -4)
    assert sorted(os.listdir(tmp_path)) == ["challenge", "challenge.hash", "response", "response.hash"]


@pytest.mark.parametrize("name", ["mnt4_753", "mnt6_753", "bls12_377"])
def test_phase2_delta_update_and_check(name):
    """Config 4: H / L query scaling by delta^-1 is bit-exact with the oracle and passes the same-ratio check
    against (delta_g2_after, delta_g2_before); a tampered element is rejected."""
    import random
    from oracle import phase2 as o2
    from snark_setup_operator_b200 import phase2 as p2
    c = get_curve(name)
    rnd = random.Random(4)
    n = 7
    pts = [c.g1.mul(c.g1.gen, rnd.randrange(1, c.Fr.p)) for _ in range(n)]
    before = ser.points_to_bytes(c.g1, pts, False)
    delta = synth.scalars_from_seed(c, synth.SEED_CONTRIB, 4)[3]
    dinv = pow(delta, -1, c.Fr.p)
    after = p2.scale_queries(name, before, n, dinv)
    assert after == o2.scale_queries(c, before, delta, False, False)
    d2_before = c.g2.mul(c.g2.gen, 31337)
    d2_after = c.g2.mul(d2_before, delta)
    db, da = ser.point_to_bytes(c.g2, d2_before, False), ser.point_to_bytes(c.g2, d2_after, False)
    p2.verify_queries(name, before, after, n, db, da, rlc_seed32=bytes(32))
    bad = bytearray(after); sz = len(after) // n
    bad[2 * sz:3 * sz] = ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 5), False)
    with pytest.raises(sso.SsoError) as e:
        p2.verify_queries(name, before, bytes(bad), n, db, da)
    assert e.value.code == -4
    # points_sum: combining partial results
    tot = p2.points_sum(name, 0, before, n)
    assert tot == ser.point_to_bytes(c.g1, c.g1.sum(pts), False)


def test_combine_and_transform_ratios(tmp_path):
    """A whole (tiny) ceremony round on the GPU: new_challenge per chunk, seeded contributions, combine, and the
    full-accumulator ratio check (transform_ratios) — all through the file-level C ABI; a tampered element is rejected."""
    import numpy as np
    name, power, cs = "bls12_377", 3, 4
    full = sso.Phase1Parameters.new_full(name, power, cs)
    p0 = sso.Phase1Parameters.new_chunk(name, 0, cs, power, cs)
    nchunks = p0.sizes()["num_chunks"]
    files, resps = [], []
    for k in range(nchunks):
        p = sso.Phase1Parameters.new_chunk(name, k, cs, power, cs)
        f = {n: str(tmp_path / ("%s_%d" % (n, k))) for n in ("challenge", "challenge.hash", "response", "response.hash", "challenge.hash2")}
        sso.new_challenge(f["challenge"], f["challenge.hash"], p)
        o = Phase1Params.new_chunk(name, k, cs, power, cs)
        ch = open(f["challenge"], "rb").read()
        assert ch == phase1.new_challenge(o)
        assert open(f["challenge.hash"], "rb").read() == phase1.calculate_hash(ch)
        sso.contribute(f["challenge"], f["challenge.hash2"], f["response"], f["response.hash"], sso.CHECK_NONZERO, 0, p, synth.SEED_CONTRIB)
        files.append(f["response"])
        resps.append(open(f["response"], "rb").read())
    combined = str(tmp_path / "combined")
    lst = str(tmp_path / "response_list")
    open(lst, "w").write("\n".join(files) + "\n")
    sso.combine(lst, combined, p0)
    o0 = Phase1Params.new_chunk(name, 0, cs, power, cs)
    assert open(combined, "rb").read() == phase1.combine(o0, resps)
    with pytest.raises(sso.SsoError) as e:                      # outputs must not exist
        sso.combine(lst, combined, p0)
    assert e.value.code == -5
    # the combined vectors are the powers of the contributor's tau on the generators
    c = get_curve(name)
    key = phase1.PrivateKey(*synth.scalars_from_seed(c, synth.SEED_CONTRIB))
    o_full = Phase1Params.new_full(name, power, cs)
    v = phase1.read_chunk(o_full, open(combined, "rb").read(), False)
    r = c.Fr.p
    assert all(c.g1.eq(P, c.g1.mul(c.g1.gen, pow(key.tau, i, r))) for i, P in enumerate(v.tau_g1))
    assert all(c.g1.eq(P, c.g1.mul(c.g1.gen, key.beta * pow(key.tau, i, r) % r)) for i, P in enumerate(v.beta_g1))
    sso.transform_ratios(combined, sso.CHECK_FULL, full, rlc_seed32=bytes(range(32)))
    sso.transform_ratios(combined, sso.CHECK_NO, full)          # fresh entropy
    assert phase1.transform_ratios(o_full, open(combined, "rb").read(), phase1.CHECK_FULL, bytes(range(32)))
    # tamper with tau_g1[5]
    mm = np.memmap(combined, dtype=np.uint8, mode="r+")
    off = 64 + 5 * 96
    mm[off:off + 96] = np.frombuffer(ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 4242), False), dtype=np.uint8)
    mm.flush(); del mm
    with pytest.raises(sso.SsoError) as e:
        sso.transform_ratios(combined, sso.CHECK_NO, full)
    assert e.value.code == -4 and "tau_g1" in e.value.message
    with pytest.raises(phase1.VerificationError, match="tau_g1"):
        phase1.transform_ratios(o_full, open(combined, "rb").read())
