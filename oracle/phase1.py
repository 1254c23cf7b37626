"""Phase-1 (powers of tau) chunk contribution and verification, restated.

Test infrastructure (see oracle/__init__.py).  Follows the call sites of the
reference — `phase1_cli::contribute` at src/bin/contribute.rs:809-824,
`phase1_cli::transform_pok_and_correctness` at src/bin/contribute.rs:966-987 and
src/bin/verify_transcript.rs:465-484, `setup_utils::calculate_hash` at
src/utils.rs:618-623 — and the upstream algorithm as recorded in SURVEY.md §3 (A),
§8a rows a2-a6 and Appendix A.3/A.4 ([UP]: `phase1`, `setup-utils` crates of
nimiq/snark-setup rev bd530da, not on disk).

Algorithm restated (per element, no batching tricks):
    generate_powers_of_tau : tau^i by an independent `pow` per index          (K1)
    batch_exp              : out[i] = (coeff * tau^i) * in[i], double-and-add (K2)
    read_batch/write_batch : arkworks byte formats, oracle/serialize.py       (K3/K4)
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

from . import serialize as ser
from .chacha import ChaChaRng, field_rand, fp_rand, group_rand
from .curves import Curve, Group
from .params import HASH_SIZE, Phase1Params


def calculate_hash(data: bytes) -> bytes:
    """setup_utils::calculate_hash = Blake2b-512, unkeyed (src/utils.rs:618-623)."""
    return hashlib.blake2b(data, digest_size=64).digest()


def blank_hash() -> bytes:
    return calculate_hash(b"")


# ---------------------------------------------------------------------------------------------
# keys
# ---------------------------------------------------------------------------------------------
@dataclass
class PrivateKey:
    tau: int
    alpha: int
    beta: int


@dataclass
class PublicKey:
    tau_g1: tuple      # (g1_s, g1_s_x)
    alpha_g1: tuple
    beta_g1: tuple
    tau_g2: object
    alpha_g2: object
    beta_g2: object

    def to_bytes(self, curve: Curve) -> bytes:
        g1, g2 = curve.g1, curve.g2
        out = b""
        for pair in (self.tau_g1, self.alpha_g1, self.beta_g1):
            out += ser.point_to_bytes(g1, pair[0], False) + ser.point_to_bytes(g1, pair[1], False)
        for p in (self.tau_g2, self.alpha_g2, self.beta_g2):
            out += ser.point_to_bytes(g2, p, False)
        return out

    @staticmethod
    def from_bytes(curve: Curve, buf: bytes) -> "PublicKey":
        g1, g2 = curve.g1, curve.g2
        s1, s2 = ser.point_size(g1, False), ser.point_size(g2, False)
        assert len(buf) == 6 * s1 + 3 * s2
        p1 = [ser.point_from_bytes(g1, buf[i * s1:(i + 1) * s1], False) for i in range(6)]
        o = 6 * s1
        p2 = [ser.point_from_bytes(g2, buf[o + i * s2:o + (i + 1) * s2], False) for i in range(3)]
        return PublicKey((p1[0], p1[1]), (p1[2], p1[3]), (p1[4], p1[5]), p2[0], p2[1], p2[2])


def hash_to_g2(curve: Curve, digest: bytes):
    """[UP] setup_utils::hash_to_g2: ChaCha20 seeded with digest[..32], then `G2::rand`.
    Highest-risk item for byte parity of the pubkey block; it does not
    touch the accumulator bytes."""
    return group_rand(curve.g2, ChaChaRng(digest[:32]))


def compute_g2_s(curve: Curve, digest: bytes, g1_s, g1_s_x, personalization: int):
    """[UP] setup_utils::compute_g2_s: Blake2b-512(personalization || digest || g1_s || g1_s_x)
    with both points in uncompressed form, then hash_to_g2."""
    h = hashlib.blake2b(digest_size=64)
    h.update(bytes([personalization]))
    h.update(digest)
    h.update(ser.point_to_bytes(curve.g1, g1_s, False))
    h.update(ser.point_to_bytes(curve.g1, g1_s_x, False))
    return hash_to_g2(curve, h.digest())


def key_generation(curve: Curve, rng: ChaChaRng, digest: bytes):
    """[UP] Phase1::key_generation: tau, alpha, beta <- Fr::rand (in that order), then for
    each of them (personalization 0, 1, 2) a proof of knowledge."""
    Fr = curve.Fr
    tau, alpha, beta = fp_rand(Fr, rng), fp_rand(Fr, rng), fp_rand(Fr, rng)

    def op(x: int, pers: int):
        g1_s = group_rand(curve.g1, rng)
        g1_s_x = curve.g1.mul(g1_s, x)
        g2_s = compute_g2_s(curve, digest, g1_s, g1_s_x, pers)
        g2_s_x = curve.g2.mul(g2_s, x)
        return (g1_s, g1_s_x), g2_s_x

    pk_tau = op(tau, 0)
    pk_alpha = op(alpha, 1)
    pk_beta = op(beta, 2)
    pub = PublicKey(pk_tau[0], pk_alpha[0], pk_beta[0], pk_tau[1], pk_alpha[1], pk_beta[1])
    return pub, PrivateKey(tau, alpha, beta)


# ---------------------------------------------------------------------------------------------
# chunk files
# ---------------------------------------------------------------------------------------------
@dataclass
class ChunkVectors:
    tau_g1: list
    tau_g2: list
    alpha_g1: list
    beta_g1: list
    beta_g2: object


def split_chunk(params: Phase1Params, buf: bytes, compressed: bool):
    """-> (hash64, raw byte slices of the five vectors)"""
    o1, o2, oa, ob, o5, end = params.offsets(compressed)
    return buf[:HASH_SIZE], (buf[o1:o2], buf[o2:oa], buf[oa:ob], buf[ob:o5], buf[o5:end])


def read_chunk(params: Phase1Params, buf: bytes, compressed: bool) -> ChunkVectors:
    c = params.curve
    _, (b1, b2, ba, bb, b5) = split_chunk(params, buf, compressed)
    return ChunkVectors(ser.points_from_bytes(c.g1, b1, compressed),
                        ser.points_from_bytes(c.g2, b2, compressed),
                        ser.points_from_bytes(c.g1, ba, compressed),
                        ser.points_from_bytes(c.g1, bb, compressed),
                        ser.point_from_bytes(c.g2, b5, compressed))


def write_chunk(params: Phase1Params, head: bytes, v: ChunkVectors, compressed: bool) -> bytes:
    c = params.curve
    assert len(head) == HASH_SIZE
    return (head + ser.points_to_bytes(c.g1, v.tau_g1, compressed)
            + ser.points_to_bytes(c.g2, v.tau_g2, compressed)
            + ser.points_to_bytes(c.g1, v.alpha_g1, compressed)
            + ser.points_to_bytes(c.g1, v.beta_g1, compressed)
            + ser.point_to_bytes(c.g2, v.beta_g2, compressed))


def new_challenge(params: Phase1Params) -> bytes:
    """phase1_cli::new_challenge (src/bin/new_setup.rs:105-109): every element = generator,
    hash slot = Blake2b-512 of the empty string."""
    c = params.curve
    v = ChunkVectors([c.g1.gen] * params.g1_count, [c.g2.gen] * params.other_count,
                     [c.g1.gen] * params.other_count, [c.g1.gen] * params.other_count, c.g2.gen)
    return write_chunk(params, blank_hash(), v, False)


# ---------------------------------------------------------------------------------------------
# contribution
# ---------------------------------------------------------------------------------------------
def generate_powers_of_tau(Fr, tau: int, start: int, end: int):
    return [pow(tau, i, Fr.p) for i in range(start, end)]


def batch_exp(G: Group, Fr, bases, exps, coeff=None):
    out = []
    for P, e in zip(bases, exps):
        k = e if coeff is None else e * coeff % Fr.p
        out.append(G.mul(P, k))
    return out


def computation(params: Phase1Params, vin: ChunkVectors, key: PrivateKey) -> ChunkVectors:
    """Phase1::computation on the parsed vectors of one chunk."""
    c = params.curve
    Fr = c.Fr
    s = params.start
    tau_pows_g1 = generate_powers_of_tau(Fr, key.tau, s, s + params.g1_count)
    tau_pows = tau_pows_g1[:params.other_count]
    return ChunkVectors(
        batch_exp(c.g1, Fr, vin.tau_g1, tau_pows_g1),
        batch_exp(c.g2, Fr, vin.tau_g2, tau_pows),
        batch_exp(c.g1, Fr, vin.alpha_g1, tau_pows, key.alpha),
        batch_exp(c.g1, Fr, vin.beta_g1, tau_pows, key.beta),
        c.g2.mul(vin.beta_g2, key.beta))


def contribute_with_key(params: Phase1Params, challenge: bytes, key: PrivateKey, pubkey_bytes: bytes) -> bytes:
    """The RNG-free core of phase1_cli::contribute: challenge (uncompressed) -> response
    (compressed || pubkey), response[0..64] = Blake2b(challenge)."""
    assert len(challenge) == params.accumulator_size, "challenge has the wrong size"
    vin = read_chunk(params, challenge, False)
    vout = computation(params, vin, key)
    out = write_chunk(params, calculate_hash(challenge), vout, True) + pubkey_bytes
    assert len(out) == params.contribution_size
    return out


def contribute(params: Phase1Params, challenge: bytes, rng: ChaChaRng):
    """phase1_cli::contribute (src/bin/contribute.rs:811-823).
    -> (response bytes, challenge_hash, response_hash)"""
    ch_hash = calculate_hash(challenge)
    pub, key = key_generation(params.curve, rng, ch_hash)
    resp = contribute_with_key(params, challenge, key, pub.to_bytes(params.curve))
    return resp, ch_hash, calculate_hash(resp)


# ---------------------------------------------------------------------------------------------
# verification building blocks
# ---------------------------------------------------------------------------------------------
CHECK_NO, CHECK_NONZERO, CHECK_FULL = 0, 1, 2            # CheckForCorrectness::{No, OnlyNonZero, Full}
SUBGROUP_AUTO, SUBGROUP_DIRECT, SUBGROUP_BATCHED, SUBGROUP_NO = 0, 1, 2, 3


class VerificationError(Exception):
    pass


def check_point(G: Group, P, check: int, subgroup: bool):
    """What `read_batch` enforces per element after parsing."""
    if check == CHECK_NO:
        return
    if P is None:
        raise VerificationError("point at infinity")
    if check == CHECK_FULL:
        if not G.on_curve(P):
            raise VerificationError("point not on curve")
        if subgroup and G.mul(P, G.r) is not None:
            raise VerificationError("point not in the prime-order subgroup")


def power_pairs_with(G: Group, v, rs):
    """power_pairs with caller-supplied randomness: (sum r_i v_i, sum r_i v_{i+1})."""
    return G.msm(v[:-1], rs), G.msm(v[1:], rs)


def merge_pairs_with(G: Group, v1, v2, rs):
    return G.msm(v1, rs), G.msm(v2, rs)


def same_ratio_dl(G1: Group, pair1, x: int) -> bool:
    """Test-only stand-in for the pairing check when the harness knows the ratio x:
    (a, b) has ratio x  <=>  b = x * a."""
    a, b = pair1
    return G1.eq(G1.mul(a, x), b)


def decompress_response(params: Phase1Params, response: bytes) -> bytes:
    """The re-encoding half of transform_pok_and_correctness: response vectors
    (compressed) -> new challenge (uncompressed) whose hash slot is Blake2b(response)."""
    body = response[:len(response) - params.public_key_size]
    v = read_chunk(params, body, True)
    return write_chunk(params, calculate_hash(response), v, False)


def verify_chunk_with_key(params: Phase1Params, challenge: bytes, response: bytes, key: PrivateKey,
                          check_out: int = CHECK_FULL, subgroup: bool = True) -> bytes:
    """Chunk verification with the pairing checks replaced by knowledge of the contributor's
    scalars (the harness made them): every output element must equal scalar * input element.
    Returns the new challenge.  Verdict-equivalent to the ratio checks of
    transform_pok_and_correctness on honest-or-corrupted outputs of a known key."""
    c = params.curve
    if len(response) != params.contribution_size:
        raise VerificationError("response has the wrong size")
    if response[:HASH_SIZE] != calculate_hash(challenge):
        raise VerificationError("hash chain broken: response does not continue the challenge")
    vin = read_chunk(params, challenge, False)
    try:
        vout = read_chunk(params, response[:len(response) - params.public_key_size], True)
    except ser.FormatError as e:
        raise VerificationError(str(e))
    for G, pts in ((c.g1, vout.tau_g1), (c.g2, vout.tau_g2), (c.g1, vout.alpha_g1), (c.g1, vout.beta_g1),
                   (c.g2, [vout.beta_g2])):
        for P in pts:
            check_point(G, P, check_out, subgroup)
    want = computation(params, vin, key)
    for G, a, b in ((c.g1, want.tau_g1, vout.tau_g1), (c.g2, want.tau_g2, vout.tau_g2),
                    (c.g1, want.alpha_g1, vout.alpha_g1), (c.g1, want.beta_g1, vout.beta_g1),
                    (c.g2, [want.beta_g2], [vout.beta_g2])):
        for P, Q in zip(a, b):
            if not G.eq(P, Q):
                raise VerificationError("ratio check failed")
    return write_chunk(params, calculate_hash(response), vout, False)


TWEAK_P1_VERIFY, TWEAK_P1_RATIOS, TWEAK_P2_VERIFY = 0x7031760000000000, 0x7031720000000000, 0x7032760000000000


def rlc_key(seed32: bytes, tweak) -> bytes:
    """ChaCha20 key of one MSM of a verification flow (csrc/curve_ops.cuh::run_msm_pairs): Blake2b-256(seed || tweak), tweak =
    (vector id, chunk index, piece offset, rank * 64 + worker slot) as four little-endian u64 — two MSMs never share their r_i
    (identical r_i on a G1 and a G2 vector would make their same_ratio comparison vacuous)."""
    import struct
    return hashlib.blake2b(seed32 + struct.pack("<4Q", *tweak), digest_size=32).digest()


def rlc_scalars(curve: Curve, seed32: bytes, n: int, tweak=None):
    """The product's reproducible random-linear-combination scalars (include/sso_b200.h,
    sso_power_pairs_dev): r_i = first 120 bits (csrc/msm.cuh RLC_BITS) of the ChaCha20(key) keystream blocks 2i, 2i+1; key = seed32 for the
    test-only primitives, rlc_key(seed32, tweak) inside the verification flows."""
    import struct
    from .chacha import chacha20_block
    if tweak is not None:
        seed32 = rlc_key(seed32, tweak)
    key = struct.unpack("<8I", seed32)
    sbits = min(curve.Fr.bits - 1, 120)                      # csrc/msm.cuh RLC_BITS
    out = []
    for i in range(n):
        words = chacha20_block(key, 2 * i)
        if sbits > 512:
            words = words + chacha20_block(key, 2 * i + 1)
        v = sum(w << (32 * j) for j, w in enumerate(words))
        out.append(v & ((1 << sbits) - 1))
    return out


def elem_policy(check_out: int, subgroup_mode: int):
    """[UP] per-element policy of Phase1::verification (the product: csrc/flows.cuh::verify_elem_check / verify_subgroup):
    response elements are always checked for zero; the membership test follows SubgroupCheckMode (No skips it) and is also
    forced by CheckForCorrectness::Full — the two knobs are independent arguments of the reference call
    (src/bin/contribute.rs:971-984)."""
    check = CHECK_NONZERO if check_out == CHECK_NO else check_out
    subgroup = subgroup_mode != SUBGROUP_NO or check_out == CHECK_FULL
    return check, subgroup


def check_point_policy(G: Group, P, check: int, subgroup: bool, uncompressed_input: bool = False):
    """csrc/kernels.cuh::body_reencode after parsing."""
    if P is None:
        if check != CHECK_NO:
            raise VerificationError("point at infinity")
        return
    if uncompressed_input and (check == CHECK_FULL or subgroup) and not G.on_curve(P):
        raise VerificationError("point not on curve")
    if subgroup and G.mul(P, G.r) is not None:
        raise VerificationError("point not in the prime-order subgroup")


def verify_chunk(params: Phase1Params, challenge: bytes, response: bytes, check_out: int = CHECK_NO, subgroup_mode: int = SUBGROUP_AUTO,
                 ratio_check: bool = True, rlc_seed32: bytes = bytes(32), check_in: int = CHECK_NO) -> bytes:
    """Phase1::verification of one chunk (or of a whole accumulator in Full mode) with real pairings (oracle/pairing.py):
    hash chain, public-key validation, the three proofs of knowledge, per-element checks, chunk-0 update checks, power-ratio
    checks on random linear combinations.  Returns the new challenge; raises VerificationError naming the failed check.
    [UP] for the exact list of checks; the product (csrc/flows.cuh::verify_chunk_host) performs the same list."""
    from .pairing import same_ratio
    c = params.curve
    g1, g2 = c.g1, c.g2
    if len(response) != params.contribution_size or len(challenge) != params.accumulator_size:
        raise VerificationError("wrong size")
    digest = calculate_hash(challenge)
    if response[:HASH_SIZE] != digest:
        raise VerificationError("hash chain broken: response does not continue the challenge")
    body = response[:len(response) - params.public_key_size]
    try:
        pub = PublicKey.from_bytes(c, response[len(body):])
        vout = read_chunk(params, body, True)
        vin = read_chunk(params, challenge, False)
    except ser.FormatError as e:
        raise VerificationError(str(e))
    check, subgroup = elem_policy(check_out, subgroup_mode)
    for G, pts in ((g1, vout.tau_g1), (g2, vout.tau_g2), (g1, vout.alpha_g1), (g1, vout.beta_g1), (g2, [vout.beta_g2])):
        for P in pts:
            check_point_policy(G, P, check, subgroup)
    if check_in != CHECK_NO:
        for G, pts in ((g1, vin.tau_g1), (g2, vin.tau_g2), (g1, vin.alpha_g1), (g1, vin.beta_g1), (g2, [vin.beta_g2])):
            for P in pts:
                check_point_policy(G, P, check_in, check_in == CHECK_FULL, uncompressed_input=True)
    for G, pts in ((g1, pub.tau_g1 + pub.alpha_g1 + pub.beta_g1), (g2, (pub.tau_g2, pub.alpha_g2, pub.beta_g2))):
        for P in pts:
            try:
                check_point_policy(G, P, CHECK_FULL, True, uncompressed_input=True)
            except VerificationError as e:
                raise VerificationError("public key: " + str(e))
    g2_s = [compute_g2_s(c, digest, pair[0], pair[1], i) for i, pair in enumerate((pub.tau_g1, pub.alpha_g1, pub.beta_g1))]
    g2_sx = [pub.tau_g2, pub.alpha_g2, pub.beta_g2]
    checks = [("proof of knowledge: tau", pub.tau_g1, (g2_s[0], g2_sx[0])),
              ("proof of knowledge: alpha", pub.alpha_g1, (g2_s[1], g2_sx[1])),
              ("proof of knowledge: beta", pub.beta_g1, (g2_s[2], g2_sx[2]))]
    if params.chunk_index == 0 and params.other_count >= 2:
        if not g1.eq(vout.tau_g1[0], g1.gen):
            raise VerificationError("tau_g1[0] is not the G1 generator")
        if not g2.eq(vout.tau_g2[0], g2.gen):
            raise VerificationError("tau_g2[0] is not the G2 generator")
        checks += [("before/after: tau_g1[1] vs tau proof", (vin.tau_g1[1], vout.tau_g1[1]), (g2_s[0], g2_sx[0])),
                   ("before/after: alpha_g1[0] vs alpha proof", (vin.alpha_g1[0], vout.alpha_g1[0]), (g2_s[1], g2_sx[1])),
                   ("before/after: beta_g1[0] vs beta proof", (vin.beta_g1[0], vout.beta_g1[0]), (g2_s[2], g2_sx[2])),
                   ("before/after: beta_g2 vs beta_g1[0]", (vin.beta_g1[0], vout.beta_g1[0]), (vin.beta_g2, vout.beta_g2))]
    if ratio_check and params.other_count >= 2:
        # a chunk past 2^power holds tau_g1 only (no G2 element): its ratios are left to transform_ratios on the combined file
        def pp(G, v, vec):
            return power_pairs_with(G, v, rlc_scalars(c, rlc_seed32, len(v) - 1, (TWEAK_P1_VERIFY | vec, params.chunk_index, 0, 0)))
        g2p = pp(g2, vout.tau_g2, 1)
        g2_exact = (vout.tau_g2[0], vout.tau_g2[1]) if params.chunk_index == 0 else g2p
        checks.append(("power ratio: tau_g1", pp(g1, vout.tau_g1, 0), g2_exact))
        checks.append(("power ratio: alpha_g1", pp(g1, vout.alpha_g1, 2), g2_exact))
        checks.append(("power ratio: beta_g1", pp(g1, vout.beta_g1, 3), g2_exact))
        if params.chunk_index == 0:
            checks.append(("power ratio: tau_g2", (vout.tau_g1[0], vout.tau_g1[1]), g2p))
    for name, p1, p2 in checks:
        if not same_ratio(c, p1, p2):
            raise VerificationError("same_ratio check failed: " + name)
    return write_chunk(params, calculate_hash(response), vout, False)


# ---------------------------------------------------------------------------------------------
# whole accumulators: combine, transform_ratios
# ---------------------------------------------------------------------------------------------
def chunk_params_of(params: Phase1Params):
    """the chunk parameters of every chunk of the ceremony `params` (any chunk of it) belongs to"""
    return [Phase1Params.new_chunk(params.curve, k, params.chunk_size, params.power, params.batch_size) for k in range(params.num_chunks)]


def combine(params: Phase1Params, responses, decompress=None) -> bytes:
    """phase1_cli::combine (src/bin/verify_transcript.rs:603-607): the vectors of all chunk responses (compressed + public key)
    concatenated into the Full-mode accumulator, uncompressed; the hash slot stays zero ([UP] Phase1::aggregation writes the
    vectors only); beta_g2 from the first chunk.  `decompress(chunk_params, response) -> five uncompressed byte strings`
    lets the C++ leg do the square roots."""
    cps = chunk_params_of(params)
    assert len(responses) == len(cps)
    vecs = [[], [], [], []]
    beta_g2 = None
    for cp, resp in zip(cps, responses):
        assert len(resp) == cp.contribution_size
        if decompress is None:
            v = read_chunk(cp, resp[:len(resp) - cp.public_key_size], True)
            c = cp.curve
            parts = (ser.points_to_bytes(c.g1, v.tau_g1, False), ser.points_to_bytes(c.g2, v.tau_g2, False),
                     ser.points_to_bytes(c.g1, v.alpha_g1, False), ser.points_to_bytes(c.g1, v.beta_g1, False),
                     ser.point_to_bytes(c.g2, v.beta_g2, False))
        else:
            parts = decompress(cp, resp)
        for i in range(4):
            vecs[i].append(parts[i])
        if beta_g2 is None:
            beta_g2 = parts[4]
    out = bytes(HASH_SIZE) + b"".join(b"".join(v) for v in vecs) + beta_g2
    full = Phase1Params.new_full(params.curve, params.power, params.batch_size)
    assert len(out) == full.accumulator_size
    return out


def transform_ratios(full: Phase1Params, combined: bytes, check_in: int = CHECK_NO, rlc_seed32: bytes = bytes(32)):
    """phase1_cli::transform_ratios (src/bin/verify_transcript.rs:811-822; [UP] Phase1::aggregate_verification): element 0 of
    tau_g1 / tau_g2 are the generators, the power ratios of the four vectors, beta_g2 against beta_g1[0].  Real pairings:
    small sizes only.  Raises VerificationError naming the failed check."""
    from .pairing import same_ratio
    c = full.curve
    g1, g2 = c.g1, c.g2
    if len(combined) != full.accumulator_size:
        raise VerificationError("wrong size")
    try:
        v = read_chunk(full, combined, False)
    except ser.FormatError as e:
        raise VerificationError(str(e))
    for G, pts in ((g1, v.tau_g1), (g2, v.tau_g2), (g1, v.alpha_g1), (g1, v.beta_g1), (g2, [v.beta_g2])):
        for P in pts:
            check_point_policy(G, P, CHECK_FULL, check_in == CHECK_FULL, uncompressed_input=True)
    if not g1.eq(v.tau_g1[0], g1.gen):
        raise VerificationError("tau_g1[0] is not the G1 generator")
    if not g2.eq(v.tau_g2[0], g2.gen):
        raise VerificationError("tau_g2[0] is not the G2 generator")

    def pp(G, vec, i):
        return power_pairs_with(G, vec, rlc_scalars(c, rlc_seed32, len(vec) - 1, (TWEAK_P1_RATIOS | i, 0, 0, 0)))
    g1p, g2p = (v.tau_g1[0], v.tau_g1[1]), (v.tau_g2[0], v.tau_g2[1])
    checks = [("power ratio: tau_g1", pp(g1, v.tau_g1, 0), g2p), ("power ratio: tau_g2", g1p, pp(g2, v.tau_g2, 1)),
              ("power ratio: alpha_g1", pp(g1, v.alpha_g1, 2), g2p), ("power ratio: beta_g1", pp(g1, v.beta_g1, 3), g2p),
              ("beta_g1[0] vs beta_g2", (v.tau_g1[0], v.beta_g1[0]), (v.tau_g2[0], v.beta_g2))]
    for name, p1, p2 in checks:
        if not same_ratio(c, p1, p2):
            raise VerificationError("same_ratio check failed: " + name)
    return True
