"""Consumes reference-held golden files when they exist (tests/golden/REFERENCE_RECIPE.md): files produced by the real
nimiq/snark-setup crates for a power-3 / chunk-4 ceremony per curve.  While tests/golden/reference/ is empty (no Rust
toolchain in the build image) every test here SKIPS — parity stays "unpinned"; with the files present the oracle (CPU)
and the CUDA path (gpu marker) are compared with them byte for byte."""
import os

import pytest

from oracle import phase1
from oracle.chacha import ChaChaRng
from oracle.curves import CURVE_NAMES
from oracle.params import Phase1Params

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "reference")
POWER, CS, BATCH = 3, 4, 4
SEED = bytes(range(32))


def _dir(name):
    d = os.path.join(REF, name)
    if not os.path.exists(os.path.join(d, "challenge_0")):
        pytest.skip("no reference-held goldens for %s (see tests/golden/REFERENCE_RECIPE.md)" % name)
    return d


def _read(d, n):
    with open(os.path.join(d, n), "rb") as fh:
        return fh.read()


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_oracle_matches_reference_files(name):
    d = _dir(name)
    o0 = Phase1Params.new_chunk(name, 0, CS, POWER, BATCH)
    resps = []
    for k in range(o0.num_chunks):
        o = Phase1Params.new_chunk(name, k, CS, POWER, BATCH)
        ch = _read(d, "challenge_%d" % k)
        assert ch == phase1.new_challenge(o), "new_challenge (generators / layout)"
        assert _read(d, "challenge_%d.hash" % k) == phase1.calculate_hash(ch)
        resp = _read(d, "response_%d" % k)
        want, _, rh = phase1.contribute(o, ch, ChaChaRng(SEED))
        body = len(resp) - o.public_key_size
        assert resp[:body] == want[:body], "accumulator bytes of the response (batch_exp, RNG -> scalars)"
        assert resp[body:] == want[body:], "public-key block (G1::rand, hash_to_g2)"
        assert _read(d, "response_%d.hash" % k) == rh
        assert _read(d, "new_challenge_%d" % k) == phase1.decompress_response(o, resp)
        resps.append(resp)
    combined = _read(d, "combined")
    assert combined == phase1.combine(o0, resps)
    of = Phase1Params.new_full(name, POWER, BATCH)
    beacon_seed = bytes.fromhex(_read(d, "beacon_seed.hex").decode().strip())
    want, _, _ = phase1.contribute(of, combined, ChaChaRng(beacon_seed))
    assert _read(d, "beacon") == want
    assert _read(d, "final") == phase1.decompress_response(of, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CURVE_NAMES)
def test_cuda_path_matches_reference_files(name, tmp_path):
    import snark_setup_operator_b200 as sso
    d = _dir(name)
    f = lambda n: str(tmp_path / n)
    p0 = sso.Phase1Parameters.new_chunk(name, 0, CS, POWER, BATCH)
    nchunks = p0.sizes()["num_chunks"]
    if name in ("mnt4_753", "mnt6_753"):                      # the reference's G2 generator, from its own round-0 challenge
        es = sso.phase1.curve_sizes(name)
        ch0 = _read(d, "challenge_0")
        g1 = ch0[64:64 + es["g1_u"]]
        g2 = ch0[64 + CS * es["g1_u"]: 64 + CS * es["g1_u"] + es["g2_u"]]
        sso.set_generators(name, g1, g2)
    try:
        files = []
        for k in range(nchunks):
            p = sso.Phase1Parameters.new_chunk(name, k, CS, POWER, BATCH)
            sso.new_challenge(f("ch%d" % k), f("ch%d.hash" % k), p)
            assert open(f("ch%d" % k), "rb").read() == _read(d, "challenge_%d" % k)
            sso.contribute(os.path.join(d, "challenge_%d" % k), f("c%d.hash" % k), f("resp%d" % k), f("resp%d.hash" % k), sso.CHECK_NO, 0, p, SEED)
            assert open(f("resp%d" % k), "rb").read() == _read(d, "response_%d" % k)
            sso.transform_pok_and_correctness(os.path.join(d, "challenge_%d" % k), f("c%d.vhash" % k), sso.CHECK_NO, os.path.join(d, "response_%d" % k),
                                              f("r%d.vhash" % k), sso.CHECK_NONZERO, f("new%d" % k), f("new%d.hash" % k), 0, True, p)
            assert open(f("new%d" % k), "rb").read() == _read(d, "new_challenge_%d" % k)
            files.append(os.path.join(d, "response_%d" % k))
        open(f("list"), "w").write("\n".join(files))
        sso.combine(f("list"), f("combined"), p0)
        assert open(f("combined"), "rb").read() == _read(d, "combined")
        pf = sso.Phase1Parameters.new_full(name, POWER, BATCH)
        beacon_seed = bytes.fromhex(_read(d, "beacon_seed.hex").decode().strip())
        sso.contribute(os.path.join(d, "combined"), f("comb.hash"), f("beacon"), f("beacon.hash"), sso.CHECK_NO, 0, pf, beacon_seed)
        assert open(f("beacon"), "rb").read() == _read(d, "beacon")
        sso.transform_pok_and_correctness(os.path.join(d, "combined"), f("cv.hash"), sso.CHECK_NO, os.path.join(d, "beacon"), f("bv.hash"),
                                          sso.CHECK_NONZERO, f("final"), f("final.hash"), 0, True, pf)
        assert open(f("final"), "rb").read() == _read(d, "final")
        sso.transform_ratios(os.path.join(d, "final"), sso.CHECK_NO, pf)
    finally:
        sso.set_generators(name, None, None)
