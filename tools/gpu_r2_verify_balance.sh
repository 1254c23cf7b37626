#!/bin/bash
# chunk verification after the stream balancing: flows tests, then the bench lines (verify seconds) of BLS12-377, BW6-761, MNT4-753
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flows.py tests/test_gpu_transcript.py -m gpu -x -q 2>&1 | tail -2
for c in bls12_377 bw6_761 mnt4_753; do
  timeout 400 python bench.py --curve $c --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> /dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$c.json")); v=d["verify"]
print("$c value %.4gM verify %.4f s runs %s in flight %.4f" % (d["value"]/1e6, v["s_per_chunk"], v["runs_s"], v["s_per_chunk_in_flight"]))
PY
done
