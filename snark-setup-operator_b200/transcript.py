"""Full-accumulator operations of `verify_transcript` / `control`: `combine` and `transform_ratios`
(reference src/bin/verify_transcript.rs:602-607, 646-653, 811-822; src/bin/control.rs:564-568, 587-591),
with the vectors sharded across the ranks of a torch.distributed group (one process per GPU).

transform_ratios is the one place on this path with an exchange step: every rank computes the
random-linear-combination pair of its contiguous shard of each vector (one-element halo for the shifted
copy), the per-rank partial points are all-gathered (NCCL over NVLink on GPUs; a few hundred bytes per
vector), every rank sums them and runs the same pairing checks.  All arithmetic is inside libsso_b200.so.
"""
from __future__ import annotations

import numpy as np

from . import phase1 as p1
from . import phase2 as p2
from ._lib import SsoError

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


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous split of the pair indices [0, n_pairs) (pair i couples elements i and i+1)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def combine(response_list, combined_filename: str, chunk_params, full_params: p1.Phase1Parameters, device=0):
    """phase1_cli::combine: concatenate the vectors of all chunk responses (compressed) into one full accumulator
    file (uncompressed here), decoding on the GPU.  `chunk_params[i]` describes response_list[i]."""
    import torch
    offs_f, counts_f, sizes_f = full_layout(full_params, False)
    total = offs_f[5]
    mm = np.memmap(combined_filename, dtype=np.uint8, mode="w+", shape=(total,))
    cursor = list(offs_f[:5])
    first = True
    for fn, cp in zip(response_list, chunk_params):
        offs_c, counts_c, sizes_c = full_layout(cp, True)
        data = np.fromfile(fn, dtype=np.uint8)
        if data.size != cp.contribution_size:
            raise SsoError(-1, "response %s has the wrong size" % fn)
        if first:
            mm[:64] = 0                                       # hash slot is filled by the caller's hash chain
        d = torch.from_numpy(data).cuda(device)
        for v in range(5):
            n = counts_c[v]
            if n == 0 or (v == 4 and not first):
                continue
            grp = 1 if v in (1, 4) else 0
            d_out = torch.empty(n * sizes_f[v], dtype=torch.uint8, device=d.device)
            p1.reencode(cp.curve, grp, d[offs_c[v]:offs_c[v + 1]], n, d_out, in_compressed=True, out_compressed=False,
                        check=p1.CHECK_NO, subgroup_check=False, device=device)
            mm[cursor[v]:cursor[v] + n * sizes_f[v]] = d_out.cpu().numpy()
            cursor[v] += n * sizes_f[v]
        first = False
    mm.flush()
    del mm


def _partial_pairs(curve, group, mm, off, elem_size, lo, hi, compressed, check, subgroup, seed, device, seg=1 << 22):
    """power_pairs over pair indices [lo, hi) of the vector stored at byte offset `off`, in segments; returns the
    list of per-segment pairs (2 uncompressed points each)."""
    import torch
    out = []
    i = lo
    while i < hi:
        j = min(hi, i + seg)
        raw = np.asarray(mm[off + i * elem_size: off + (j + 1) * elem_size])       # elements i .. j inclusive (halo)
        d = torch.from_numpy(raw.copy()).cuda(device)
        out.append(p1.power_pairs(curve, group, d, j - i + 1, in_compressed=compressed, check=check, subgroup_check=subgroup,
                                  seed32=seed, device=device))
        i = j
    return out


def transform_ratios(combined_filename: str, check_input: int, full_params: p1.Phase1Parameters, compressed=False, device=0,
                     rlc_seed32=None, subgroup_check=False):
    """phase1_cli::transform_ratios: power-ratio verification of a full accumulator.  Uses every rank of the
    default torch.distributed group if one is initialised.  Raises SsoError(code -4) on rejection."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    curve = full_params.curve
    es = p1.curve_sizes(curve)
    offs, counts, sizes = full_layout(full_params, compressed)
    mm = np.memmap(combined_filename, dtype=np.uint8, mode="r")
    if mm.size < offs[5]:
        raise SsoError(-1, "combined file is too small for these parameters")
    usz = (es["g1_u"], es["g2_u"])
    pairs = []
    for v in range(4):
        grp = VEC_GROUPS[v]
        lo, hi = shard_range(counts[v] - 1, rank, world)
        parts = _partial_pairs(curve, grp, mm, offs[v], sizes[v], lo, hi, compressed, check_input, subgroup_check, rlc_seed32, device) \
            if hi > lo else []
        # fold this rank's segments, then exchange the per-rank partial points
        ident = bytes(usz[grp] - 1) + b"\x40"
        flat_a = b"".join(p[:usz[grp]] for p in parts) or ident
        flat_b = b"".join(p[usz[grp]:] for p in parts) or ident
        a = p2.points_sum(curve, grp, flat_a, max(1, len(parts)), device=device)
        b = p2.points_sum(curve, grp, flat_b, max(1, len(parts)), device=device)
        if world > 1:
            mine = torch.frombuffer(bytearray(a + b), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                mine = mine.cuda(device)
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)                   # the single collective of the path
            blobs = [bytes(g.cpu().numpy().tobytes()) for g in gathered]
            a = p2.points_sum(curve, grp, b"".join(x[:usz[grp]] for x in blobs), world, device=device)
            b = p2.points_sum(curve, grp, b"".join(x[usz[grp]:] for x in blobs), world, device=device)
        pairs.append((a, b))

    def elem(v, i, uncompressed_size, grp):
        raw = bytes(mm[offs[v] + i * sizes[v]: offs[v] + (i + 1) * sizes[v]])
        if not compressed:
            return raw
        import torch as _t
        d_in = _t.frombuffer(bytearray(raw), dtype=_t.uint8).cuda(device)
        d_out = _t.empty(uncompressed_size, dtype=_t.uint8, device=d_in.device)
        p1.reencode(curve, grp, d_in, 1, d_out, in_compressed=True, out_compressed=False, check=p1.CHECK_NO, subgroup_check=False,
                    device=device)
        return bytes(d_out.cpu().numpy().tobytes())

    g1_0, g1_1 = elem(0, 0, usz[0], 0), elem(0, 1, usz[0], 0)
    g2_0, g2_1 = elem(1, 0, usz[1], 1), elem(1, 1, usz[1], 1)
    beta_g1_0 = elem(3, 0, usz[0], 0)
    beta_g2 = elem(4, 0, usz[1], 1)
    names = ["power ratio: tau_g1", "power ratio: tau_g2", "power ratio: alpha_g1", "power ratio: beta_g1", "beta_g1[0] vs beta_g2"]
    checks = [(pairs[0][0], pairs[0][1], g2_0, g2_1), (g1_0, g1_1, pairs[1][0], pairs[1][1]),
              (pairs[2][0], pairs[2][1], g2_0, g2_1), (pairs[3][0], pairs[3][1], g2_0, g2_1),
              (g1_0, beta_g1_0, g2_0, beta_g2)]
    verdicts = p1.same_ratio(curve, checks, device=device)
    for nm, ok in zip(names, verdicts):
        if not ok:
            raise SsoError(-4, "same_ratio check failed: " + nm)
    return True
