#!/usr/bin/env python3
"""Diagnostic: repeat contribute_dev and the IMAD probe, logging wall time, NVML clocks and power."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import snark_setup_operator_b200 as sso
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
def state():
    return "sm=%dMHz mem=%dMHz P=%.0fW T=%dC reasons=%x" % (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
        pynvml.nvmlDeviceGetPowerUsage(h)/1000, pynvml.nvmlDeviceGetTemperature(h, 0), pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
cs = 1 << 16
p = sso.Phase1Parameters.new_chunk("bls12_377", 1, cs, 20, cs)
sz = p.sizes()
d_gen = torch.empty(sz["accumulator_size"], dtype=torch.uint8, device="cuda")
sso.new_challenge_dev(p, d_gen)
d_resp = torch.zeros(sz["contribution_size"], dtype=torch.uint8, device="cuda")
k = (0x1234567890abcdef1234567890abcdef1234567890abcdef, 0xfedcba9876543210fedcba9876543210fedcba98765432, 0x1111111111111111222222222222222233333333333333)
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
print("start", state(), flush=True)
for it in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sso.contribute_dev(p, d_gen, d_resp, *k, check=0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pk = sso.imad_peak(0)
    t2 = time.perf_counter()
    print("iter %2d contribute %.1f ms  imad %.2f TMAC/s (%.0f ms)  %s" % (it, (t1-t0)*1e3, pk/1e12, (t2-t1)*1e3, state()), flush=True)
# now G1-only vector calls back to back
g1u = 96
d_in = d_gen[64:64 + cs*g1u]
d_out = torch.zeros(cs*48, dtype=torch.uint8, device="cuda")
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sso.batch_exp("bls12_377", 0, d_in, cs, 65536, k[0], k[1], d_out)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("g1 batch_exp %.1f ms %s" % ((t1-t0)*1e3, state()), flush=True)
