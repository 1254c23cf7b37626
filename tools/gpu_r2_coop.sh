#!/bin/bash
# A/B of the warp-cooperative G2 bodies (coop.cuh; SSO_COOP_G2=0 selects the one-thread-per-element bodies): parity tests with
# the cooperative bodies (the default), then the contribute bench per curve in both modes.
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/${TAG}_parity.log
tail -3 gpurun_out/${TAG}_parity.log
for mode in 1 0; do
  SSO_COOP_G2=$mode timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_bls12_377_coop$mode.json 2> gpurun_out/${TAG}_bench_bls12_377_coop$mode.err
done
for c in mnt4_753 mnt6_753; do
  SSO_COOP_G2=1 timeout 300 python bench.py --curve $c --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${c}_coop1.json 2> gpurun_out/${TAG}_bench_${c}_coop1.err
done
python - <<PY
import json, glob
for fn in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.load(open(fn))
    except Exception as e:
        print(fn, "FAILED", e); continue
    rf = d.get("roofline") or {}
    print(fn.split("/")[-1], "%.4g" % d["value"], d["unit"], "ms/step %.2f" % d["ms_per_step"], "frac %.4f" % rf.get("frac", 0), "e2e", (d.get("e2e") or {}).get("value"))
PY
