#!/bin/bash
# after the width-4 NAF ladders (membership tests, cofactor clearing) and the norm-map square root of Fq3:
# parity + flows tests, then the bench lines (contribute value, chunk verify seconds) per curve
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flows.py -m gpu -x -q 2>&1 | tail -2
for c in bls12_377 bw6_761 mnt4_753 mnt6_753; do
  timeout 400 python bench.py --curve $c --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> /dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$c.json")); v=d["verify"]; e=d["e2e"]
print("$c value %.4gM frac %.4f e2e %.4gM seeded_call %.1f ms verify %.4f s runs %s in flight %.4f" % (d["value"]/1e6, d["roofline"]["frac"], e["value"]/1e6, e["seeded_call"]["ms_per_step"], v["s_per_chunk"], v["runs_s"], v["s_per_chunk_in_flight"]))
PY
done
bash tools/gpu_verify_ncu.sh ${TAG} mnt6_753 > gpurun_out/${TAG}_verify_kernels_mnt6.txt 2>&1; rm -f gpurun_out/${TAG}_verify_*.csv; head -9 gpurun_out/${TAG}_verify_kernels_mnt6.txt
