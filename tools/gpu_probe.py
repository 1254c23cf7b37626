#!/usr/bin/env python3
"""Exploratory GPU measurements (not the bench): multiply-accumulate peaks and per-curve
device-resident contribute timings.  Usage: python tools/gpu_probe.py [curve:chunk_log ...]"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import snark_setup_operator_b200 as sso  # noqa: E402


def gen_challenge_image(curve, n_g1, n_g2, acc_size, offs, sizes):
    """all-generator accumulator built on the device by decompressing one generator (no oracle)."""
    raise NotImplementedError


def main():
    specs = sys.argv[1:] or ["bls12_377:16"]
    res = {"device": torch.cuda.get_device_name(0)}
    out = ctypes.c_double(0)
    for v, nm in ((0, "mad.wide.u32"), (1, "mad.lo.cc+madc.hi.cc chain"), (2, "mad.lo.u32")):
        sso._lib.call("sso_imad_peak", 0, v, ctypes.byref(out))
        res["imad_peak_%d" % v] = {"what": nm, "macs_per_s": out.value}
        print("imad probe", v, nm, "%.3e MAC/s" % out.value, flush=True)
    from oracle import phase1, synth
    from oracle.curves import get_curve
    from oracle.params import Phase1Params
    for spec in specs:
        name, clog = spec.split(":")
        clog = int(clog)
        cs = 1 << clog
        c = get_curve(name)
        o = Phase1Params.new_chunk(name, 1, cs, clog + 4, cs)
        p = sso.Phase1Parameters.new_chunk(name, 1, cs, clog + 4, cs)
        t0 = time.time()
        gen = phase1.new_challenge(o)
        d_gen = torch.frombuffer(bytearray(gen), dtype=torch.uint8).cuda()
        k1 = synth.scalars_from_seed(c, synth.SEED_PREV)
        k2 = synth.contributor_key(c)
        d_r = torch.zeros(o.contribution_size, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        print(name, "setup %.1fs" % (time.time() - t0), flush=True)
        times = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.time()
            sso.contribute_dev(p, d_gen, d_r, k1[0], k1[1], k1[2])
            torch.cuda.synchronize()
            times.append(time.time() - t0)
        npts = o.g1_count + 3 * o.other_count + 1
        best = min(times[1:])
        res[spec] = {"points": npts, "wall_s": times, "points_per_s": npts / best}
        print(name, clog, "points", npts, "times", ["%.4f" % t for t in times], "-> %.0f points/s" % (npts / best), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
