"""World-size-2 `gloo` test of the cooperative transform_ratios algebra, on CPU.

The product's exchange step runs inside libsso_b200.so (csrc/stream.cuh: pieces dealt round-robin over the ranks, one
partial pair per piece with its own Blake2b-derived ChaCha20 key, per-rank fold, ONE all-gather, point addition) and needs
GPUs; here the same piece / owner / tweak logic is driven with the oracle standing in for the device arithmetic, over a real
torch.distributed group: the gathered-and-folded pair must equal the pair computed piece by piece on a single rank, every
rank must see the same gathered bytes, and the result must satisfy the power ratio.  The GPU-side NCCL branch is covered by
tests/test_gpu_multi.py (needs two devices) and the kept hardware logs under profiles/."""
import os

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp


def _partial(c, G, pts, lo, cnt, seed, vec_id, rank_slot):
    from oracle import phase1
    rs = phase1.rlc_scalars(c, seed, cnt - 1, (phase1.TWEAK_P1_RATIOS | vec_id, 0, lo, rank_slot))
    return phase1.power_pairs_with(G, pts[lo:lo + cnt], rs)


def _worker(rank, world, port, n, piece, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import serialize as ser, synth
    from oracle.curves import get_curve
    from snark_setup_operator_b200.transcript import owner, pieces
    c = get_curve("bls12_377")
    G = c.g1
    s = synth.scalars_from_seed(c, synth.SEED_PREV)[0]
    pts = synth._powers_points(G, G.gen, s, c.Fr.p, 0, n)
    seed = bytes(range(32))
    a = b = None
    for i, (lo, cnt) in enumerate(pieces(n, piece, True)):
        if owner(i, world) != rank:
            continue
        pa, pb = _partial(c, G, pts, lo, cnt, seed, 0, rank * 64)
        a, b = G.add(a, pa), G.add(b, pb)
    mine = torch.frombuffer(bytearray(ser.point_to_bytes(G, a, False) + ser.point_to_bytes(G, b, False)), dtype=torch.uint8)
    got = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)                                  # the single collective of the path
    ret[rank] = [bytes(g.numpy().tobytes()) for g in got]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,piece", [(2, 8), (9, 8), (37, 8), (64, 16)])
def test_cooperative_power_pairs_over_gloo(n, piece):
    from oracle import serialize as ser, synth
    from oracle.curves import get_curve
    from snark_setup_operator_b200.transcript import owner, pieces
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (n * 7 + piece) % 97
    mp.spawn(_worker, args=(world, port, n, piece, ret), nprocs=world, join=True)
    assert ret[0] == ret[1]                                   # every rank sees the same gathered partials
    c = get_curve("bls12_377")
    G = c.g1
    usz = 2 * G.F.nbytes
    a = b = None
    for blob in ret[0]:
        a = G.add(a, ser.point_from_bytes(G, blob[:usz], False))
        b = G.add(b, ser.point_from_bytes(G, blob[usz:], False))
    # single-rank reference: the same pieces with the keys their owners used
    s = synth.scalars_from_seed(c, synth.SEED_PREV)[0]
    pts = synth._powers_points(G, G.gen, s, c.Fr.p, 0, n)
    ps = pieces(n, piece, True)
    # the pieces cover every pair index exactly once, with one element of halo
    assert ps[0][0] == 0 and all(ps[i][0] + ps[i][1] - 1 == ps[i + 1][0] for i in range(len(ps) - 1)) and ps[-1][0] + ps[-1][1] == n
    wa = wb = None
    for i, (lo, cnt) in enumerate(ps):
        pa, pb = _partial(c, G, pts, lo, cnt, bytes(range(32)), 0, owner(i, world) * 64)
        wa, wb = G.add(wa, pa), G.add(wb, pb)
    assert G.eq(a, wa) and G.eq(b, wb)
    assert G.eq(G.mul(a, s), b)                               # the combination keeps the power ratio


def test_chunk_to_rank_assignment():
    """contribute path: chunks are independent; rank r of N takes chunks r, r+N, ... (no collective)."""
    n_chunks, world = 32, 8
    seen = sorted(k for r in range(world) for k in range(r, n_chunks, world))
    assert seen == list(range(n_chunks))
