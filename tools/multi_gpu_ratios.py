#!/usr/bin/env python3
"""Multi-GPU full-accumulator ratio verification (transform_ratios) — run under torchrun:
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_ratios.py [curve] [power]
Rank 0 builds a synthetic combined accumulator (the product's own contribute on the full all-generator
accumulator); every rank then verifies its shard, partial MSM points are all-gathered over NCCL."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import snark_setup_operator_b200 as sso
from snark_setup_operator_b200 import transcript

curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_377"
power = int(sys.argv[2]) if len(sys.argv) > 2 else 14
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
full = sso.Phase1Parameters.new_full(curve, power, 1 << power)
path = "/tmp/sso_combined_%s_%d" % (curve, power)
if rank == 0:
    d_gen = torch.empty(full.accumulator_size, dtype=torch.uint8, device="cuda")
    sso.new_challenge_dev(full, d_gen, device=local)
    resp = torch.zeros(full.contribution_size, dtype=torch.uint8, device="cuda")
    sso.contribute_dev(full, d_gen, resp, 0x1234567, 0x89abcdef, 0x13579bdf, check=0, device=local)
    open(path + ".resp", "wb").write(resp.cpu().numpy().tobytes())
    transcript.combine([path + ".resp"], path, [full], full, device=local)
if world > 1:
    warm = torch.zeros(1, device="cuda")
    dist.all_reduce(warm)                      # NCCL communicator set-up is not part of the measurement
    dist.barrier()
sso.imad_peak(0, local)                        # CUDA context / module load
torch.cuda.synchronize()
t0 = time.perf_counter()
ok = transcript.transform_ratios(path, sso.CHECK_NO, full, device=local)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = t.item()
if rank == 0:
    sz = full.sizes()
    print({"curve": curve, "power": power, "points": sz["g1_count"] + 3 * sz["other_count"] + 1, "n_gpus": world, "verify_s": dt, "ok": ok})
if world > 1:
    dist.destroy_process_group()
