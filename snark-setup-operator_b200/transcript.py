"""Full-accumulator operations of `verify_transcript` / `control`: `combine` and `transform_ratios`
(reference src/bin/verify_transcript.rs:602-607, 646-653, 811-822; src/bin/control.rs:564-568, 587-591).

Both live behind the C ABI now (`sso_p1_combine_file`, `sso_p1_verify_ratios_file`: streamed in `batch_size` pieces over
the devices of the process or the ranks of a process group, one NCCL all-gather of the partial MSM results inside the
library); this module keeps the host-side helpers the tests and tools share: the layout arithmetic and the contiguous
shard split used to reason about the pieces.
"""
from __future__ import annotations

from . import phase1 as p1
from .phase1 import combine, transform_ratios  # noqa: F401  (re-exported: the reference-shaped calls)

VEC_GROUPS = (0, 1, 0, 0)          # tauG1, tauG2, alphaG1, betaG1
VEC_NAMES = ("tau_g1", "tau_g2", "alpha_g1", "beta_g1")


def full_layout(params: p1.Phase1Parameters, compressed: bool):
    """byte offsets of (tauG1, tauG2, alphaG1, betaG1, betaG2, end) and element counts of a chunk/full file"""
    sz = params.sizes()
    es = p1.curve_sizes(params.curve)
    g1 = es["g1_c"] if compressed else es["g1_u"]
    g2 = es["g2_c"] if compressed else es["g2_u"]
    counts = (sz["g1_count"], sz["other_count"], sz["other_count"], sz["other_count"], 1)
    sizes = (g1, g2, g1, g1, g2)
    offs = [64]
    for n, s in zip(counts, sizes):
        offs.append(offs[-1] + n * s)
    return offs, counts, sizes


def pieces(n: int, piece: int, pairs: bool):
    """The pieces the library cuts a vector of n elements into (csrc/stream.cuh::stream_reencode): (first element, count).
    With pairs (power_pairs) piece k covers pair indices [lo, hi) and therefore elements [lo, hi] — one element of halo."""
    out = []
    if pairs and n >= 2:
        lo = 0
        while lo < n - 1:
            hi = min(lo + piece, n - 1)
            out.append((lo, hi - lo + 1))
            lo = hi
    else:
        lo = 0
        while lo < n:
            out.append((lo, min(piece, n - lo)))
            lo += piece
    return out


def owner(piece_index: int, world: int) -> int:
    """rank that processes piece `piece_index` of a cooperative call (round-robin)"""
    return piece_index % world
