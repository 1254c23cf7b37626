#!/bin/bash
# Round-2 single-GPU measurement set: default bench (BLS12-377), per-curve benches, phase-2 workload, verify_transcript workload.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
for c in bw6_761 mnt4_753 mnt6_753; do
  timeout 400 python bench.py --curve $c --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_$c.json 2> gpurun_out/${TAG}_bench_n1_$c.err
done
for c in mnt4_753 mnt6_753; do
  for q in 19 20; do
    timeout 300 python bench.py --workload phase2 --curve $c --query-log $q --steps 3 > gpurun_out/${TAG}_phase2_${c}_q$q.json 2> gpurun_out/${TAG}_phase2_${c}_q$q.err
  done
done
timeout 600 python bench.py --workload verify_transcript --curve bw6_761 --power 18 --chunk-log 15 --steps 1 --warmup 1 > gpurun_out/${TAG}_vt_bw6_p18_n1.json 2> gpurun_out/${TAG}_vt_bw6_p18_n1.err
python - <<PY
import json, glob
for fn in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d = json.load(open(fn))
    except Exception as e:
        print(fn, "FAILED", e); continue
    extra = ""
    if d.get("verify"): extra += " verify %.3f s (in flight %.3f)" % (d["verify"]["s_per_chunk"], d["verify"]["s_per_chunk_in_flight"])
    if d.get("phases_s"): extra += " phases " + json.dumps({k: round(v, 2) for k, v in d["phases_s"].items()})
    rf = d.get("roofline") or {}
    print(fn.split("/")[-1], d["metric"], "%.4g" % d["value"], d["unit"], "frac", rf.get("frac"), "e2e", (d.get("e2e") or {}).get("value"), extra)
PY
