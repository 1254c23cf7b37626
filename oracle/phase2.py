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
