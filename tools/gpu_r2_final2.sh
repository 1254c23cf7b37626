#!/bin/bash
# Final single-GPU measurement set of round 2 (everything lands under gpurun_out/ as small text / JSON; the .ncu-rep files are
# summarised on the box and removed: the merge back is capped at 64 MiB).
#   tools/gpu_r2_final2.sh TAG [tests]
TAG=${1:-r2z}
mkdir -p gpurun_out
if [ "$2" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
  tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 20 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
for c in bw6_761 mnt4_753 mnt6_753; do
  timeout 600 python bench.py --curve $c --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_$c.json 2> gpurun_out/${TAG}_bench_n1_$c.err
done
for spec in "mnt4_753 20" "mnt6_753 20" "bls12_377 22"; do
  set -- $spec
  timeout 300 python bench.py --workload phase2 --curve $1 --query-log $2 --steps 3 > gpurun_out/${TAG}_phase2_$1_q$2.json 2> gpurun_out/${TAG}_phase2_$1_q$2.err
done
# launch list of the default bench command (share of the step per kernel)
SSO_BENCH_NOSAMPLER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
# full captures of the chunk kernel: BLS12-377 at the bench size, MNT4-753 (cooperative G2) at chunk 2^14
capture() {   # name title bench-args...
  local name=$1 title=$2; shift 2
  SSO_BENCH_NOVERIFY=1 SSO_BENCH_QUICK=1 SSO_BENCH_NOSAMPLER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_batch_exp_chunk \
    --launch-skip 3 --launch-count 1 -o gpurun_out/${TAG}_$name -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}_ncu.log 2>&1
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > /tmp/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page source --csv > /tmp/${name}_source.csv 2>/dev/null
  python tools/ncu_raw_summary.py /tmp/${name}_raw.csv "$title" > gpurun_out/${TAG}_${name}_summary.csv
  python tools/ncu_sass_regions.py /tmp/${name}_source.csv 20 > gpurun_out/${TAG}_${name}_sass_regions.txt
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
capture ncu_bls12_377 "k_batch_exp_chunk, BLS12-377 2^20 / chunk 2^16 (196608 G1 + 65537 G2 points), final round-2 build: ncu --set full --clock-control none --import-source on, bench.py --steps 1 --warmup 3 (quick mode), 4th launch"
capture ncu_mnt4_753 "k_batch_exp_chunk, MNT4-753 chunk 2^14 (49152 G1 + 16385 G2 points, cooperative G2 body), final round-2 build: ncu --set full, 4th launch" --curve mnt4_753 --chunk-log 14
bash tools/gpu_verify_ncu.sh ${TAG} bls12_377 mnt4_753 > gpurun_out/${TAG}_verify_kernels.txt 2>&1
rm -f gpurun_out/${TAG}_verify_*.csv
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
cat gpurun_out/${TAG}_verify_kernels.txt | head -30
du -sh gpurun_out
