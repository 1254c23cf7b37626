#!/usr/bin/env python3
"""Sweep of host workers x chunks for the chunks-in-flight calls on the bench configuration (ms per chunk)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import snark_setup_operator_b200 as sso
cs = 1 << 16
p = sso.Phase1Parameters.new_chunk("bls12_377", 1, cs, 20, cs)
d_gen = torch.empty(p.accumulator_size, dtype=torch.uint8, device="cuda")
sso.new_challenge_dev(p, d_gen)
ch = torch.empty(p.accumulator_size, dtype=torch.uint8).pin_memory()
ch.copy_(d_gen)
NMAX = 24
resps = [torch.empty(p.contribution_size, dtype=torch.uint8).pin_memory() for _ in range(NMAX)]
seed = bytes(range(32))
for seeded in (False, True):
    for workers in (2, 3, 4, 6, 8):
        row = []
        for n in (6, 8, 12, 24):
            def run():
                if seeded:
                    sso.contribute_seeded_many_buf([p] * n, [ch] * n, resps[:n], seed, check=0, host_threads=workers)
                else:
                    sso.contribute_many_buf([p] * n, [ch] * n, resps[:n], 3, 5, 7, pubkey=None, check=0, host_threads=workers)
            run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            row.append((time.perf_counter() - t0) * 1e3 / n)
        print("seeded=%d workers=%d  ms/chunk for n=6,8,12,24: %s" % (seeded, workers, " ".join("%.1f" % x for x in row)), flush=True)
