#!/usr/bin/env python3
"""Host timeline (SSO_TRACE=1) of one full chunk contribution (sso_p1_contribute_seeded_buf) on the bench configuration."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import snark_setup_operator_b200 as sso
curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_377"
cs = 1 << 16
p = sso.Phase1Parameters.new_chunk(curve, 1, cs, 20, cs)
d_gen = torch.empty(p.accumulator_size, dtype=torch.uint8, device="cuda")
sso.new_challenge_dev(p, d_gen)
ch = torch.empty(p.accumulator_size, dtype=torch.uint8).pin_memory()
ch.copy_(d_gen)
resp = torch.empty(p.contribution_size, dtype=torch.uint8).pin_memory()
sso.contribute_seeded_buf(p, ch, resp, bytes(range(32)), check=0)
for i in range(2):
    sys.stderr.write("==== call %d ====\n" % (i + 2)); sys.stderr.flush()
    t0 = time.perf_counter()
    sso.contribute_seeded_buf(p, ch, resp, bytes(range(32)), check=0)
    sys.stderr.write("wall %.2f ms\n" % ((time.perf_counter() - t0) * 1e3))
