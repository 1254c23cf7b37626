#!/bin/bash
# Fourth measurement of round 2: unreduced operands also in the one-thread-per-element towers (BLS12-377 G2, the verification
# ladders); parity + flows tests, contribute bench per curve (quick mode), full bench line (with chunk verification) for MNT4-753.
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flows.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -2 gpurun_out/${TAG}_tests.log
run() {   # name curve steps
  local name=$1 curve=$2 steps=$3
  SSO_BENCH_NOVERIFY=1 SSO_BENCH_QUICK=1 timeout 300 python bench.py --curve $curve --steps $steps --warmup 3 --no-cpu-baseline \
      > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$name.json"))
    ks={k["kernel"]:round(k["ms_total"]/max(1,k["launches"]),3) for k in d["roofline"]["kernels"]}
    print("$name", "ms/step %.3f" % d["ms_per_step"], "value %.4fM" % (d["value"]/1e6), "frac %.4f" % d["roofline"]["frac"], ks, d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name FAILED", e)
PY
}
run bls  bls12_377 10
run bw6  bw6_761 5
run mnt4 mnt4_753 3
run mnt6 mnt6_753 3
timeout 400 python bench.py --curve mnt4_753 --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_full_mnt4.json 2> gpurun_out/${TAG}_full_mnt4.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_full_mnt4.json"))
print("mnt4 full: value %.4fM e2e %.4fM verify %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, json.dumps(d.get("verify"))[:300]))
PY
