"""World-size-2 `gloo` test of the host-side sharding logic used by transform_ratios and the benchmark:
the contiguous split with a one-element halo covers every pair index exactly once, and the all_gather of
the per-rank partial results delivers every rank's bytes to every rank."""
import os

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_pairs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snark_setup_operator_b200.transcript import shard_range
    lo, hi = shard_range(n_pairs, rank, world)
    # stand-in for the partial pair: (sum of pair indices, count) — additive like the real partial points
    mine = torch.tensor([sum(range(lo, hi)), hi - lo, lo, hi], dtype=torch.int64)
    got = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    ret[rank] = [g.tolist() for g in got]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [1, 2, 7, 1 << 16])
def test_shards_cover_all_pairs_and_gather(n_pairs):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (n_pairs % 97)
    mp.spawn(_worker, args=(world, port, n_pairs, ret), nprocs=world, join=True)
    assert ret[0] == ret[1]                                   # every rank sees the same gathered partials
    parts = ret[0]
    assert sum(p[1] for p in parts) == n_pairs
    assert sum(p[0] for p in parts) == n_pairs * (n_pairs - 1) // 2
    assert parts[0][2] == 0 and parts[-1][3] == n_pairs and parts[0][3] == parts[1][2]   # contiguous, no gap, no overlap


def test_chunk_to_rank_assignment():
    """contribute path: chunks are independent; rank r of N takes chunks r, r+N, ... (no collective)."""
    n_chunks, world = 32, 8
    seen = sorted(k for r in range(world) for k in range(r, n_chunks, world))
    assert seen == list(range(n_chunks))
