"""Host-side mirror of the phase-2 path over the C ABI: the delta update of the Groth16 H / L queries
(`phase2_cli::contribute::<P>`, reference src/bin/contribute.rs:826-839) and its same-ratio check
(`phase2_cli::verify::<P>`, src/bin/contribute.rs:989-1008).  Only the scaling / verification cores are
bound; the MPCParameters container (ark-groth16 ProvingKey layout) stays with the host application."""
from __future__ import annotations

import ctypes

from ._lib import call
from .phase1 import CHECK_NO, curve_id, curve_sizes, scalar_bytes


def scale_queries(curve, queries, n: int, delta_inv: int, in_compressed=False, out_compressed=False, check=CHECK_NO,
                  device=0, out=None):
    """h_query / l_query (n serialized G1 points: bytes, numpy or a pinned torch tensor) -> the same vector multiplied by
    delta^-1; returns bytes, or fills the writable buffer `out` when given."""
    from .phase1 import _host_ptr
    s = curve_sizes(curve)
    in_ptr, in_len, k1 = _host_ptr(queries)
    out_len = n * (s["g1_c"] if out_compressed else s["g1_u"])
    buf = ctypes.create_string_buffer(out_len) if out is None else None
    out_ptr, out_len2, k2 = (ctypes.addressof(buf), out_len, buf) if out is None else _host_ptr(out)
    call("sso_p2_scale_queries_buf", curve_id(curve), in_ptr, in_len, out_ptr, out_len2, n, scalar_bytes(curve, delta_inv),
         int(in_compressed), int(out_compressed), check, device)
    del k1, k2
    return buf.raw if out is None else out


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
