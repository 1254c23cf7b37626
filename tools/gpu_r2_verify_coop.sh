#!/bin/bash
# membership test of the MNT G2 groups through the cooperative kernel: tests, then the bench lines of MNT4-753 / MNT6-753
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flows.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_transcript.py -m gpu -x -q -k "verdicts or subgroup_knob or public_key or tail_chunk" 2>&1 | tail -2
for c in mnt4_753 mnt6_753; do
  timeout 400 python bench.py --curve $c --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> /dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$c.json")); v=d["verify"]; e=d["e2e"]
print("$c value %.4gM frac %.4f e2e %.4gM verify %.4f s runs %s in flight %.4f" % (d["value"]/1e6, d["roofline"]["frac"], e["value"]/1e6, v["s_per_chunk"], v["runs_s"], v["s_per_chunk_in_flight"]))
PY
done
SSO_COOP_G2=0 timeout 300 python - <<'PY'
# the same verification with the membership test inside the decoding kernel (A/B in one process image)
import json, subprocess, sys
PY
