"""Host-side mirror of the phase-2 path over the C ABI: the delta update of the Groth16 H / L queries
(`phase2_cli::contribute::<P>`, reference src/bin/contribute.rs:826-839) and its same-ratio check
(`phase2_cli::verify::<P>`, src/bin/contribute.rs:989-1008).  Only the scaling / verification cores are
bound; the MPCParameters container (ark-groth16 ProvingKey layout) stays with the host application."""
from __future__ import annotations

import ctypes

from ._lib import call
from .phase1 import CHECK_NO, curve_id, curve_sizes, scalar_bytes


def scale_queries(curve, queries: bytes, n: int, delta_inv: int, in_compressed=False, out_compressed=False, check=CHECK_NO,
                  device=0) -> bytes:
    """h_query / l_query (n serialized G1 points) -> the same vector multiplied by delta^-1."""
    s = curve_sizes(curve)
    out = ctypes.create_string_buffer(n * (s["g1_c"] if out_compressed else s["g1_u"]))
    call("sso_p2_scale_queries_buf", curve_id(curve), queries, len(queries), out, len(out), n, scalar_bytes(curve, delta_inv),
         int(in_compressed), int(out_compressed), check, device)
    return out.raw


def verify_queries(curve, before: bytes, after: bytes, n: int, delta_g2_before: bytes, delta_g2_after: bytes,
                   before_compressed=False, after_compressed=False, check=CHECK_NO, subgroup_check=False, rlc_seed32=None, device=0):
    """Raises SsoError(code -4) unless same_ratio(merge_pairs(before, after), (delta_g2_after, delta_g2_before))."""
    call("sso_p2_verify_queries_buf", curve_id(curve), before, len(before), after, len(after), n, int(before_compressed),
         int(after_compressed), delta_g2_before, delta_g2_after, check, int(subgroup_check), rlc_seed32, device)


def points_sum(curve, group: int, points: bytes, n: int, device=0) -> bytes:
    """Sum of n uncompressed points (combining all-gathered per-GPU partial MSM results)."""
    s = curve_sizes(curve)
    out = ctypes.create_string_buffer(s["g1_u"] if group == 0 else s["g2_u"])
    call("sso_points_sum", curve_id(curve), group, points, n, out, len(out), device)
    return out.raw
