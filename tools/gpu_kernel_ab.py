#!/usr/bin/env python3
"""Times the G1 / G2 batch_exp launches in isolation (one vector per call, one stream) for the library
named by SSO_B200_LIB.  Usage: python tools/gpu_kernel_ab.py [curve] [log2 n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import snark_setup_operator_b200 as sso
from snark_setup_operator_b200.phase1 import curve_sizes
curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_377"
n = 1 << int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
es = curve_sizes(curve)
p = sso.Phase1Parameters.new_chunk(curve, 1, n, 24, n)
sz = p.sizes()
d_gen = torch.empty(sz["accumulator_size"], dtype=torch.uint8, device="cuda")
sso.new_challenge_dev(p, d_gen)
k = (0x1234567890abcdef1234567890abcdef1234567890abcdef, 0xfedcba9876543210fedcba9876543210fedcba98765432)
res = {}
for g, nm in ((0, "g1"), (1, "g2")):
    off = 64 if g == 0 else 64 + n * es["g1_u"]
    d_in = d_gen[off: off + n * es[nm + "_u"]]
    d_out = torch.zeros(n * es[nm + "_c"], dtype=torch.uint8, device="cuda")
    # randomise the points first so that the inputs are not all the generator
    sso.batch_exp(curve, g, d_in, n, 5, k[1], None, d_out)
    d_in2 = torch.zeros(n * es[nm + "_u"], dtype=torch.uint8, device="cuda")
    sso.reencode(curve, g, d_out, n, d_in2, check=0, subgroup_check=False)
    sso.profile_reset(); sso.profile_enable(True)
    for it in range(3):
        sso.batch_exp(curve, g, d_in2, n, 65536, k[0], k[1], d_out)
    pr = sso.profile_read()
    res[nm] = (pr["batch_exp_" + nm]["ms"] / 3, pr["normalize_" + nm]["ms"] / 3)
    sso.profile_enable(False)
print(os.environ.get("SSO_B200_LIB", "default"), curve, n, " ".join("%s exp %.3f ms norm %.3f ms" % (nm, a, b) for nm, (a, b) in res.items()))
