#!/usr/bin/env python3
"""Timings of the verification building blocks on the GPU: power_pairs per vector and same_ratio batches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import snark_setup_operator_b200 as sso
from snark_setup_operator_b200.phase1 import curve_sizes
for curve, clog in (("bls12_377", 16), ("bw6_761", 14), ("mnt4_753", 14), ("mnt6_753", 14)):
    n = 1 << clog
    es = curve_sizes(curve)
    p = sso.Phase1Parameters.new_chunk(curve, 1, n, 24, n)
    sz = p.sizes()
    d_gen = torch.empty(sz["accumulator_size"], dtype=torch.uint8, device="cuda")
    sso.new_challenge_dev(p, d_gen)
    d_resp = torch.zeros(sz["contribution_size"], dtype=torch.uint8, device="cuda")
    k = (0x1234567890abcdef1234567890abcdef1234567890abcdef, 0xfedcba9876543210fedcba9876543210fedcba98765432, 0x3333333333333333444444444444444455555555555555)
    sso.contribute_dev(p, d_gen, d_resp, *k, check=0)
    torch.cuda.synchronize()
    out = []
    off = 64
    for g, nm in ((0, "g1"), (1, "g2")):
        d_in = d_resp[off: off + n * es[nm + "_c"]]
        off += n * es[nm + "_c"]
        for chk, sub in ((0, False), (2, True)):
            t0 = time.perf_counter()
            pair = sso.power_pairs(curve, g, d_in, n, in_compressed=True, check=chk, subgroup_check=sub)
            torch.cuda.synchronize()
            out.append("%s power_pairs(check=%d,subgroup=%d) %.1f ms" % (nm, chk, sub, (time.perf_counter() - t0) * 1e3))
    # same_ratio batch of 8 checks built from the generator pair
    g1u, g2u = es["g1_u"], es["g2_u"]
    gen = bytes(d_gen[:64 + g1u].cpu().numpy().tobytes())[64:]
    gen2 = bytes(d_gen[64 + n * g1u: 64 + n * g1u + g2u].cpu().numpy().tobytes())
    checks = [(gen, gen, gen2, gen2)] * 8
    for reps in range(2):
        t0 = time.perf_counter()
        v = sso.same_ratio(curve, checks)
        dt = (time.perf_counter() - t0) * 1e3
    out.append("same_ratio x8 %.1f ms %s" % (dt, all(v)))
    print(curve, "n=2^%d:" % clog, "; ".join(out), flush=True)
