"""Deterministic synthetic accumulators (SURVEY.md §8d "Synthetic inputs").

Test infrastructure (see oracle/__init__.py).  A chunk as it looks after a previous
contributor with scalars (s, a, b) = first three `Fr::rand` draws of
ChaCha20(seed = 0x5e * 32):  tauG1[j] = s^i G1, tauG2[j] = s^i G2, alphaG1 = a s^i G1,
betaG1 = b s^i G1, betaG2 = b G2 with i = chunk_index * chunk_size + j; hash slot =
Blake2b-512 of the empty string.  The contributor under test draws (tau, alpha, beta)
from ChaCha20(seed = bytes(range(32))).
"""
from __future__ import annotations

from .chacha import ChaChaRng, fp_rand
from .params import Phase1Params
from .phase1 import ChunkVectors, PrivateKey, blank_hash, write_chunk

SEED_PREV = bytes([0x5E] * 32)
SEED_CONTRIB = bytes(range(32))


def scalars_from_seed(curve, seed: bytes, n: int = 3):
    rng = ChaChaRng(seed)
    return [fp_rand(curve.Fr, rng) for _ in range(n)]


def contributor_key(curve) -> PrivateKey:
    t, a, b = scalars_from_seed(curve, SEED_CONTRIB)
    return PrivateKey(t, a, b)


def _powers_points(G, base, s: int, r: int, start: int, n: int):
    """[s^start * base, s^(start+1) * base, ...] by repeated scalar multiplication."""
    if n == 0:
        return []
    P = G.mul(base, pow(s, start, r))
    out = [P]
    for _ in range(n - 1):
        P = G.mul(P, s)
        out.append(P)
    return out


def synthetic_vectors(params: Phase1Params, seed: bytes = SEED_PREV) -> ChunkVectors:
    c = params.curve
    r = c.Fr.p
    s, a, b = scalars_from_seed(c, seed)
    st, g1n, on = params.start, params.g1_count, params.other_count
    tau_g1 = _powers_points(c.g1, c.g1.gen, s, r, st, g1n)
    tau_g2 = _powers_points(c.g2, c.g2.gen, s, r, st, on)
    alpha_g1 = [c.g1.mul(P, a) for P in tau_g1[:on]]
    beta_g1 = [c.g1.mul(P, b) for P in tau_g1[:on]]
    return ChunkVectors(tau_g1, tau_g2, alpha_g1, beta_g1, c.g2.mul(c.g2.gen, b))


def synthetic_challenge(params: Phase1Params, seed: bytes = SEED_PREV) -> bytes:
    return write_chunk(params, blank_hash(), synthetic_vectors(params, seed), False)
