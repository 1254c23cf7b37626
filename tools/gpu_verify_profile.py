#!/usr/bin/env python3
"""Per-kernel-kind device time (CUDA events) of one chunk verification on the bench configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import snark_setup_operator_b200 as sso
curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_377"
cs = 1 << 16
p = sso.Phase1Parameters.new_chunk(curve, 1, cs, 20, cs)
d_gen = torch.empty(p.accumulator_size, dtype=torch.uint8, device="cuda")
sso.new_challenge_dev(p, d_gen)
ch = d_gen.cpu().numpy()
resp = bytearray(p.contribution_size)
sso.contribute_seeded_buf(p, ch, resp, bytes(range(32)), check=0)
new = bytearray(p.accumulator_size)
sso.verify_chunk_buf(p, ch, bytes(resp), new)
sso.profile_reset()
sso.profile_enable(2)
sso.verify_chunk_buf(p, ch, bytes(resp), new)
sso.profile_enable(False)
for k, v in sso.profile_read().items():
    if v["launches"]:
        print("%-18s launches %3d  %9.3f ms" % (k, v["launches"], v["ms"]))
