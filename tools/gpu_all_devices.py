#!/usr/bin/env python3
"""One process, all GPUs of the box (device = -1): the chunks-in-flight work queue spread over every visible device.
Prints ms per chunk for the bench configuration and checks the responses against the single-device call; then a small
MNT6-753 batch (its block inversion tree needs > 48 KB of dynamic shared memory: the per-device function attribute)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import snark_setup_operator_b200 as sso
ndev = torch.cuda.device_count()
cs = 1 << 16
p = sso.Phase1Parameters.new_chunk("bls12_377", 1, cs, 20, cs)
d_gen = torch.empty(p.accumulator_size, dtype=torch.uint8, device="cuda:0")
sso.new_challenge_dev(p, d_gen)
ch = torch.empty(p.accumulator_size, dtype=torch.uint8).pin_memory()
ch.copy_(d_gen)
n = 6 * ndev
resps = [torch.empty(p.contribution_size, dtype=torch.uint8).pin_memory() for _ in range(n)]
ref = torch.empty(p.contribution_size, dtype=torch.uint8).pin_memory()
seed = bytes(range(32))
sso.contribute_seeded_buf(p, ch, ref, seed, check=0, device=0)
for seeded in (False, True):
    def run():
        if seeded:
            sso.contribute_seeded_many_buf([p] * n, [ch] * n, resps, seed, check=0, host_threads=6 * ndev, device=-1)
        else:
            sso.contribute_many_buf([p] * n, [ch] * n, resps, 3, 5, 7, check=0, host_threads=3 * ndev, device=-1)
    run()
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    npts = 262145
    print("devices=%d seeded=%d chunks=%d  %.2f ms per chunk  %.1f M points/s" % (ndev, seeded, n, dt * 1e3 / n, n * npts / dt / 1e6), flush=True)
    if seeded:
        print("identical to the single-device call:", all(torch.equal(r, ref) for r in resps), flush=True)
# MNT6-753, small chunk, every device
q = sso.Phase1Parameters.new_chunk("mnt6_753", 1, 64, 8, 64)
g = torch.empty(q.accumulator_size, dtype=torch.uint8, device="cuda:0")
sso.new_challenge_dev(q, g)
h = g.cpu().numpy()
outs = [bytearray(q.contribution_size) for _ in range(2 * ndev)]
sso.contribute_seeded_many_buf([q] * len(outs), [h] * len(outs), outs, seed, check=0, host_threads=2 * ndev, device=-1)
print("mnt6_753 on all devices identical:", all(o == outs[0] for o in outs), flush=True)
