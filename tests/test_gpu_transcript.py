"""verify_transcript-level runs on the GPU through the file-level C ABI (reference src/bin/verify_transcript.rs:293-569,
602-607, 675-696, 745-776, 811-822): new_challenge -> chunk contributions -> chunk verifications -> combine -> beacon
contribution in Full mode -> its verification -> transform_ratios, every output compared byte for byte with the oracle
(Python big-int leg for the keys and the tiny cases, C++ leg for the bulk arithmetic), plus the parity holes of round 1:
tail chunks, all four curves / both groups for the rejects, check_input, the subgroup knob, the public key, the RLC seed."""
import hashlib
import os

import pytest

import snark_setup_operator_b200 as sso
from oracle import cport, phase1, serialize as ser, synth
from oracle.chacha import ChaChaRng
from oracle.curves import CURVE_NAMES, _some_point, get_curve
from oracle.params import Phase1Params

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

BEACON = hashlib.blake2s(b"beacon value of the test ceremony").digest()          # 32 bytes, as verify_transcript.rs:161-166 enforces


def _oracle_contribute(o, challenge, seed, threads=0):
    """phase1_cli::contribute by the oracle: keys by the Python leg, vectors by the C++ leg"""
    digest = phase1.calculate_hash(challenge)
    pub, key = phase1.key_generation(o.curve, ChaChaRng(seed), digest)
    return cport.contribute_with_key(o, challenge, key, pub.to_bytes(o.curve), threads=threads)


def _run_transcript(tmp_path, name, power, cs_log, batch, devices=None):
    cs = 1 << cs_log
    p0 = sso.Phase1Parameters.new_chunk(name, 0, cs, power, batch)
    o0 = Phase1Params.new_chunk(name, 0, cs, power, batch)
    nchunks = o0.num_chunks
    f = lambda n: str(tmp_path / n)
    resp_files, resps = [], []
    for k in range(nchunks):
        p = sso.Phase1Parameters.new_chunk(name, k, cs, power, batch)
        o = Phase1Params.new_chunk(name, k, cs, power, batch)
        # round 0: the initial challenge of the chunk
        sso.new_challenge(f("ch%d" % k), f("ch%d.hash" % k), p)
        ch = open(f("ch%d" % k), "rb").read()
        assert ch == cport.new_challenge(o)
        assert open(f("ch%d.hash" % k), "rb").read() == phase1.calculate_hash(ch)
        # the contribution (every chunk with the same seed-derived key, as one contributor does)
        sso.contribute(f("ch%d" % k), f("ch%d.hash2" % k), f("resp%d" % k), f("resp%d.hash" % k), sso.CHECK_NONZERO, 0, p, synth.SEED_CONTRIB)
        resp = open(f("resp%d" % k), "rb").read()
        assert resp == _oracle_contribute(o, ch, synth.SEED_CONTRIB), "chunk %d response" % k
        assert open(f("resp%d.hash" % k), "rb").read() == phase1.calculate_hash(resp)
        # its verification (operator defaults: check_input No, non-zero outputs, subgroup Auto, ratio check on)
        sso.transform_pok_and_correctness(f("ch%d" % k), f("ch%d.vhash" % k), sso.CHECK_NO, f("resp%d" % k), f("resp%d.vhash" % k),
                                          sso.CHECK_NONZERO, f("new%d" % k), f("new%d.hash" % k), 0, True, p)
        new = open(f("new%d" % k), "rb").read()
        assert new == cport.decompress_response(o, resp, check=0, subgroup=False)
        assert open(f("new%d.hash" % k), "rb").read() == phase1.calculate_hash(new)
        assert open(f("ch%d.vhash" % k), "rb").read() == phase1.calculate_hash(ch)
        resp_files.append(f("resp%d" % k))
        resps.append(resp)
    # aggregation
    open(f("response_list"), "w").write("\n".join(resp_files) + "\n")
    sso.combine(f("response_list"), f("combined"), p0, devices=devices)
    combined = open(f("combined"), "rb").read()
    assert combined == cport.combine(o0, resps)
    # the random beacon: phase1_cli::contribute on the combined accumulator, Full mode, rng seeded with the beacon hash
    pf = sso.Phase1Parameters.new_full(name, power, batch)
    of = Phase1Params.new_full(name, power, batch)
    sso.contribute(f("combined"), f("combined.hash"), f("beacon"), f("beacon.hash"), sso.CHECK_NONZERO, 0, pf, BEACON)
    beacon = open(f("beacon"), "rb").read()
    assert beacon == _oracle_contribute(of, combined, BEACON)
    assert open(f("combined.hash"), "rb").read() == phase1.calculate_hash(combined)
    assert open(f("beacon.hash"), "rb").read() == phase1.calculate_hash(beacon)
    # verification of the beacon contribution (Full mode: proofs of knowledge, first-element updates, all power ratios)
    sso.transform_pok_and_correctness(f("combined"), f("combined.vhash"), sso.CHECK_NO, f("beacon"), f("beacon.vhash"), sso.CHECK_NONZERO,
                                      f("final"), f("final.hash"), 0, True, pf)
    final = open(f("final"), "rb").read()
    assert final == cport.decompress_response(of, beacon, check=0, subgroup=False)
    assert open(f("final.hash"), "rb").read() == phase1.calculate_hash(final)
    # consistency of the whole accumulator
    sso.transform_ratios(f("final"), sso.CHECK_NO, pf, devices=devices)
    sso.transform_ratios(f("combined"), sso.CHECK_NO, pf, devices=devices)
    return f, o0, of, final


def test_transcript_bls12_377_power14(tmp_path):
    """power 14, chunks of 2^12 (8 chunks), pieces of 2^10: every file of the run equals the oracle's; a wrong element in the
    G1-only tail of the final accumulator is rejected by transform_ratios."""
    import numpy as np
    f, o0, of, final = _run_transcript(tmp_path, "bls12_377", 14, 12, 1 << 10)
    pf = sso.Phase1Parameters.new_full("bls12_377", 14, 1 << 10)
    c = of.curve
    mm = np.memmap(f("final"), dtype=np.uint8, mode="r+")
    off = 64 + ((1 << 14) + 4321) * 96                              # tau_g1 beyond 2^power
    mm[off:off + 96] = np.frombuffer(ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 987654321), False), dtype=np.uint8)
    mm.flush(); del mm
    with pytest.raises(sso.SsoError) as e:
        sso.transform_ratios(f("final"), sso.CHECK_NO, pf)
    assert e.value.code == -4 and "tau_g1" in e.value.message


@pytest.mark.parametrize("name", ["bw6_761", "mnt4_753", "mnt6_753"])
def test_transcript_small_other_curves(tmp_path, name):
    """the same run at power 5 / chunk 2^3 / pieces of 4 on the 24-limb curves"""
    _run_transcript(tmp_path, name, 5, 3, 4)


def test_tail_chunk_wrong_scalar_is_caught_by_transform_ratios(tmp_path):
    """A chunk past 2^power holds tau_g1 only: a wrong scalar in its middle cannot be seen by the chunk verification (no G2
    element to compare with) and is rejected by transform_ratios on the combined accumulator — documented in
    include/sso_b200.h; [UP] upstream's per-chunk verification leaves the ratios to aggregate_verification."""
    name, power, cs = "bls12_377", 3, 4
    c = get_curve(name)
    p0 = sso.Phase1Parameters.new_chunk(name, 0, cs, power, cs)
    files = []
    for k in range(4):
        p = sso.Phase1Parameters.new_chunk(name, k, cs, power, cs)
        o = Phase1Params.new_chunk(name, k, cs, power, cs)
        ch = phase1.new_challenge(o)
        resp = bytearray(o.contribution_size)
        sso.contribute_seeded_buf(p, ch, resp, synth.SEED_CONTRIB)
        if k == 2:                                                  # first tail chunk (start = 8 = 2^power)
            assert o.other_count == 0
            resp[64 + 48: 64 + 96] = ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 424242), True)
            resp_hash_fix = bytes(resp)                             # the hash chain is recomputed by the verifier from the files
            new = bytearray(o.accumulator_size)
            sso.verify_chunk_buf(p, ch, resp_hash_fix, new, ratio_check=True, rlc_seed32=bytes(range(32)))    # accepted
        fn = str(tmp_path / ("r%d" % k))
        open(fn, "wb").write(resp)
        files.append(fn)
    open(str(tmp_path / "list"), "w").write("\n".join(files))
    sso.combine(str(tmp_path / "list"), str(tmp_path / "combined"), p0)
    with pytest.raises(sso.SsoError) as e:
        sso.transform_ratios(str(tmp_path / "combined"), sso.CHECK_NO, sso.Phase1Parameters.new_full(name, power, cs))
    assert e.value.code == -4 and "tau_g1" in e.value.message


def _contribution(name, k, cs=4, power=3):
    o = Phase1Params.new_chunk(name, k, cs, power, cs)
    p = sso.Phase1Parameters.new_chunk(name, k, cs, power, cs)
    ch = synth.synthetic_challenge(o)
    resp = bytearray(o.contribution_size)
    sso.contribute_seeded_buf(p, ch, resp, synth.SEED_CONTRIB)
    return o, p, ch, bytes(resp)


def _non_subgroup_point(G):
    """a curve point outside the order-r subgroup (None when the group has cofactor 1)"""
    for start in (11, 12, 13, 14, 15):
        P = _some_point(G, start)
        if G.mul(P, G.r) is not None:
            return P
    return None


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_verify_verdicts_match_oracle_on_every_curve(name):
    """accept / reject verdicts of sso_p1_verify_chunk_buf against oracle.phase1.verify_chunk (real pairings) on all four
    curves and both groups: wrong scalar (G1 and G2), non-subgroup point (G1 where the cofactor is not 1, G2), infinity,
    tampered public key."""
    o, p, ch, resp = _contribution(name, 0)
    c = o.curve
    oc = o.offsets(True)
    s1, s2 = c.g1.F.nbytes, c.g2.F.nbytes
    new_ch = bytearray(o.accumulator_size)
    seed = bytes(range(32))
    sso.verify_chunk_buf(p, ch, resp, new_ch, rlc_seed32=seed)
    assert phase1.verify_chunk(o, ch, resp, rlc_seed32=seed) == bytes(new_ch)

    def both_reject(bad, match, oracle_too=True, **kw):
        with pytest.raises(sso.SsoError) as e:
            sso.verify_chunk_buf(p, ch, bytes(bad), new_ch, rlc_seed32=seed, **kw)
        assert e.value.code == -4 and match in e.value.message, e.value.message
        if oracle_too:
            with pytest.raises(phase1.VerificationError):
                phase1.verify_chunk(o, ch, bytes(bad), rlc_seed32=seed)

    bad = bytearray(resp); bad[oc[0] + 2 * s1: oc[0] + 3 * s1] = ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 12345), True)
    both_reject(bad, "power ratio: tau_g1")
    bad = bytearray(resp); bad[oc[1] + 2 * s2: oc[1] + 3 * s2] = ser.point_to_bytes(c.g2, c.g2.mul(c.g2.gen, 12345), True)
    both_reject(bad, "power ratio")
    for G, vec, sz in ((c.g1, 3, s1), (c.g2, 1, s2)):
        P = _non_subgroup_point(G)
        if P is None:
            continue                                                # MNT G1: cofactor 1
        bad = bytearray(resp); bad[oc[vec] + sz: oc[vec] + 2 * sz] = ser.point_to_bytes(G, P, True)
        both_reject(bad, "subgroup")
        # SubgroupCheckMode::No lets the membership test go; the ratio check still sees a wrong element
        with pytest.raises(sso.SsoError) as e:
            sso.verify_chunk_buf(p, ch, bytes(bad), new_ch, rlc_seed32=seed, check_output=sso.CHECK_NONZERO, subgroup_check_mode=sso.phase1.SUBGROUP_NO)
        assert "subgroup" not in e.value.message
    bad = bytearray(resp); bad[oc[1] + s2: oc[1] + 2 * s2] = ser.point_to_bytes(c.g2, None, True)
    both_reject(bad, "infinity")
    bad = bytearray(resp); pk0 = oc[5]
    bad[pk0 + 2 * s1: pk0 + 4 * s1] = ser.point_to_bytes(c.g1, c.g1.mul(c.g1.gen, 99), False)
    both_reject(bad, "proof of knowledge: tau", oracle_too=(name == "bls12_377"))


def test_subgroup_knob_and_check_input():
    """the membership test follows subgroup_check_mode, not check_output (round-1 finding); check_input is honoured"""
    o, p, ch, resp = _contribution("bls12_377", 1)
    c = o.curve
    oc, ou = o.offsets(True), o.offsets(False)
    new_ch = bytearray(o.accumulator_size)
    P = _non_subgroup_point(c.g1)
    bad = bytearray(resp); bad[oc[0] + 48: oc[0] + 96] = ser.point_to_bytes(c.g1, P, True)
    for check_output in (sso.CHECK_NO, sso.CHECK_NONZERO, sso.CHECK_FULL):                      # Auto: always tested
        with pytest.raises(sso.SsoError) as e:
            sso.verify_chunk_buf(p, ch, bytes(bad), new_ch, check_output=check_output, ratio_check=False)
        assert e.value.code == -4 and "subgroup" in e.value.message
    sso.verify_chunk_buf(p, ch, bytes(bad), new_ch, check_output=sso.CHECK_NONZERO, subgroup_check_mode=sso.phase1.SUBGROUP_NO, ratio_check=False)
    with pytest.raises(sso.SsoError):                                                            # Full forces it
        sso.verify_chunk_buf(p, ch, bytes(bad), new_ch, check_output=sso.CHECK_FULL, subgroup_check_mode=sso.phase1.SUBGROUP_NO, ratio_check=False)
    # check_input: a challenge element off the subgroup — the response built from it is what an attacker would send
    bad_ch = bytearray(ch); bad_ch[ou[0] + 96: ou[0] + 192] = ser.point_to_bytes(c.g1, P, False)
    bad_ch = bytes(bad_ch)
    resp2 = bytearray(o.contribution_size)
    sso.contribute_seeded_buf(p, bad_ch, resp2, synth.SEED_CONTRIB, check=sso.CHECK_NONZERO)    # default contribute check: accepted
    with pytest.raises(sso.SsoError) as e:                                                       # --force-correctness-checks
        sso.contribute_seeded_buf(p, bad_ch, resp2, synth.SEED_CONTRIB, check=sso.CHECK_FULL)
    assert "subgroup" in e.value.message
    with pytest.raises(sso.SsoError) as e:
        sso.verify_chunk_buf(p, bad_ch, bytes(resp2), new_ch, check_input=sso.CHECK_FULL, subgroup_check_mode=sso.phase1.SUBGROUP_NO, ratio_check=False)
    assert e.value.code == -4 and "challenge" in e.value.message


@pytest.mark.parametrize("name", ["bls12_377", "mnt4_753"])
def test_public_key_points_are_validated(name):
    """infinity or a torsion point in the public key would make the proof-of-knowledge pairings vacuous (advisor finding)"""
    o, p, ch, resp = _contribution(name, 1)
    c = o.curve
    pk0 = o.offsets(True)[5]
    s1, s2 = 2 * c.g1.F.nbytes, 2 * c.g2.F.nbytes
    new_ch = bytearray(o.accumulator_size)
    bad = bytearray(resp)
    bad[pk0: pk0 + s1] = ser.point_to_bytes(c.g1, None, False)              # tau proof: g1_s = g1_s_x = O
    bad[pk0 + s1: pk0 + 2 * s1] = ser.point_to_bytes(c.g1, None, False)
    with pytest.raises(sso.SsoError) as e:
        sso.verify_chunk_buf(p, ch, bytes(bad), new_ch)
    assert e.value.code == -4 and "public key" in e.value.message
    with pytest.raises(phase1.VerificationError, match="public key"):
        phase1.verify_chunk(o, ch, bytes(bad))
    bad = bytearray(resp)
    bad[pk0 + 6 * s1: pk0 + 6 * s1 + s2] = ser.point_to_bytes(c.g2, _non_subgroup_point(c.g2), False)
    with pytest.raises(sso.SsoError) as e:
        sso.verify_chunk_buf(p, ch, bytes(bad), new_ch)
    assert e.value.code == -4 and "public key" in e.value.message


def test_seeded_rlc_does_not_share_scalars_between_msms():
    """With a caller-supplied RLC seed every MSM still gets its own scalars: tau_g1[i] = x_i G1, tau_g2[i] = x_i G2 with
    arbitrary x_i pass a G1-vs-G2 power-ratio comparison built from IDENTICAL r_i (advisor finding); here they are rejected."""
    import random
    o, p, ch, resp = _contribution("bls12_377", 1)
    c = o.curve
    oc = o.offsets(True)
    rnd = random.Random(9)
    xs = [rnd.randrange(1, c.Fr.p) for _ in range(o.other_count)]
    bad = bytearray(resp)
    bad[oc[0]:oc[1]] = ser.points_to_bytes(c.g1, [c.g1.mul(c.g1.gen, x) for x in xs], True)
    bad[oc[1]:oc[2]] = ser.points_to_bytes(c.g2, [c.g2.mul(c.g2.gen, x) for x in xs], True)
    new_ch = bytearray(o.accumulator_size)
    with pytest.raises(sso.SsoError) as e:
        sso.verify_chunk_buf(p, ch, bytes(bad), new_ch, rlc_seed32=bytes(range(32)))
    assert e.value.code == -4 and "power ratio" in e.value.message
    with pytest.raises(phase1.VerificationError, match="power ratio"):
        phase1.verify_chunk(o, ch, bytes(bad), rlc_seed32=bytes(range(32)))


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_new_challenge_bytes_and_generator_override(tmp_path, name):
    """phase1_cli::new_challenge byte parity with the oracle on every curve (chunk and Full parameters), and the generator
    hand-over: after sso_p1_set_generators the initial accumulator and the chunk-0 generator check use the caller's points."""
    c = get_curve(name)
    for tag, p, o in (("chunk", sso.Phase1Parameters.new_chunk(name, 1, 4, 3, 4), Phase1Params.new_chunk(name, 1, 4, 3, 4)),
                      ("full", sso.Phase1Parameters.new_full(name, 3, 4), Phase1Params.new_full(name, 3, 4))):
        fn, hn = str(tmp_path / (tag + ".ch")), str(tmp_path / (tag + ".hash"))
        sso.new_challenge(fn, hn, p)
        got = open(fn, "rb").read()
        assert got == phase1.new_challenge(o)
        assert open(hn, "rb").read() == phase1.calculate_hash(got)
    g1b, g2b = c.g1.mul(c.g1.gen, 7), c.g2.mul(c.g2.gen, 7)
    try:
        sso.set_generators(name, ser.point_to_bytes(c.g1, g1b, False), ser.point_to_bytes(c.g2, g2b, False))
        p, o = sso.Phase1Parameters.new_chunk(name, 0, 4, 3, 4), Phase1Params.new_chunk(name, 0, 4, 3, 4)
        fn, hn = str(tmp_path / "ovr.ch"), str(tmp_path / "ovr.hash")
        sso.new_challenge(fn, hn, p)
        got = open(fn, "rb").read()
        want = (phase1.blank_hash() + ser.point_to_bytes(c.g1, g1b, False) * o.g1_count + ser.point_to_bytes(c.g2, g2b, False) * o.other_count
                + ser.point_to_bytes(c.g1, g1b, False) * (2 * o.other_count) + ser.point_to_bytes(c.g2, g2b, False))
        assert got == want
        # chunk 0 of a ceremony on those generators verifies; on the built-in ones it would fail the generator check
        resp = bytearray(o.contribution_size)
        sso.contribute_seeded_buf(p, got, resp, synth.SEED_CONTRIB)
        new = bytearray(o.accumulator_size)
        sso.verify_chunk_buf(p, got, bytes(resp), new)
        with pytest.raises(sso.SsoError):                               # not a subgroup point: refused
            sso.set_generators(name, ser.point_to_bytes(c.g1, g1b, False), ser.point_to_bytes(c.g2, _non_subgroup_point(c.g2), False))
    finally:
        sso.set_generators(name, None, None)
    with pytest.raises(sso.SsoError) as e:
        sso.verify_chunk_buf(p, got, bytes(resp), new)
    assert "generator" in e.value.message


@pytest.mark.parametrize("name,group", [("bls12_377", 0), ("bls12_377", 1), ("bw6_761", 0)])
def test_msm_at_2_16_against_closed_form(name, group):
    """power_pairs at n = 2^16 (the bench chunk size; window width chosen by msm_window_bits) against a closed form: the
    vector is v_i = s^i G, so (sum r_i v_i, sum r_i v_{i+1}) = ((sum r_i s^i) G, (sum r_i s^(i+1)) G) with the reproducible r_i
    — two oracle scalar multiplications instead of 2^16 point additions."""
    c = get_curve(name)
    G = (c.g1, c.g2)[group]
    r = c.Fr.p
    n = 1 << 16
    s = synth.scalars_from_seed(c, synth.SEED_PREV)[0]
    es = sso.phase1.curve_sizes(name)
    usz = es["g1_u" if group == 0 else "g2_u"]
    gen = ser.point_to_bytes(G, G.gen, False)
    d_in = torch.frombuffer(bytearray(gen * n), dtype=torch.uint8).cuda()
    d_c = torch.empty(n * usz // 2, dtype=torch.uint8, device="cuda")
    sso.batch_exp(name, group, d_in, n, 0, s, None, d_c)                    # v_i = s^i G (compressed)
    seed = bytes(range(32))
    pair = sso.power_pairs(name, group, d_c, n, in_compressed=True, check=sso.CHECK_FULL, subgroup_check=False, seed32=seed)
    rs = phase1.rlc_scalars(c, seed, n - 1)
    ka = kb = 0
    sp = 1
    for ri in rs:
        ka = (ka + ri * sp) % r
        sp = sp * s % r
        kb = (kb + ri * sp) % r
    assert pair[:usz] == ser.point_to_bytes(G, G.mul(G.gen, ka), False)
    assert pair[usz:] == ser.point_to_bytes(G, G.mul(G.gen, kb), False)


@pytest.mark.parametrize("name", ["mnt4_753", "mnt6_753"])
def test_config3_full_size_spot_checked_by_the_cpp_oracle(name):
    """BASELINE config 3 at its named size (2^20 powers, chunk 2^16): the GPU response of a full chunk, with sampled elements
    of every vector recomputed one by one by the C++ oracle leg (per-index pow + double-and-add), and the decompress-compress
    identity on the whole response."""
    import random
    power, cs, k = 20, 1 << 16, 1
    p = sso.Phase1Parameters.new_chunk(name, k, cs, power, cs)
    o = Phase1Params.new_chunk(name, k, cs, power, cs)
    c = o.curve
    prev = phase1.PrivateKey(*synth.scalars_from_seed(c, synth.SEED_PREV))
    key = synth.contributor_key(c)
    # the challenge: a previous contribution applied to the all-generator accumulator, on the GPU
    d_gen = torch.empty(p.accumulator_size, dtype=torch.uint8, device="cuda")
    sso.new_challenge_dev(p, d_gen)
    gen = d_gen.cpu().numpy().tobytes()
    r1 = bytearray(o.contribution_size)
    sso.contribute_buf(p, gen, r1, prev.tau, prev.alpha, prev.beta, pubkey=bytes(o.public_key_size), check=sso.CHECK_NO)
    ch = bytearray(o.accumulator_size)
    ou, oc = o.offsets(False), o.offsets(True)
    counts = (o.g1_count, o.other_count, o.other_count, o.other_count, 1)
    groups = (0, 1, 0, 0, 1)
    ch[:64] = gen[:64]
    for v in range(5):
        d_c = torch.frombuffer(bytearray(r1[oc[v]:oc[v + 1]]), dtype=torch.uint8).cuda()
        d_u = torch.empty(ou[v + 1] - ou[v], dtype=torch.uint8, device="cuda")
        sso.reencode(name, groups[v], d_c, counts[v], d_u, check=sso.CHECK_NO, subgroup_check=False)
        ch[ou[v]:ou[v + 1]] = d_u.cpu().numpy().tobytes()
    ch = bytes(ch)
    resp = bytearray(o.contribution_size)
    sso.contribute_buf(p, ch, resp, key.tau, key.alpha, key.beta, pubkey=bytes(o.public_key_size), check=sso.CHECK_NONZERO)
    rnd = random.Random(5)
    coeffs = (None, None, key.alpha, key.beta)
    for v in range(4):
        G = (c.g1, c.g2)[groups[v]]
        usz, csz = 2 * G.F.nbytes, G.F.nbytes
        for j in [0, 1, counts[v] - 1] + [rnd.randrange(counts[v]) for _ in range(5)]:
            one = ch[ou[v] + j * usz: ou[v] + (j + 1) * usz]
            want = cport.batch_exp(c, groups[v], one, 1, o.start + j, key.tau, coeffs[v], threads=1)
            assert bytes(resp[oc[v] + j * csz: oc[v] + (j + 1) * csz]) == want, (v, j)
    want = cport.batch_exp(c, 1, ch[ou[4]:ou[5]], 1, 0, 1, key.beta, mode=1, threads=1)
    assert bytes(resp[oc[4]:oc[5]]) == want
