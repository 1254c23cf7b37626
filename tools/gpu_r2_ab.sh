#!/bin/bash
# A/B of library variants on the bench workload (device-resident value + roofline pass only).
#   tools/gpu_r2_ab.sh TAG variant1 variant2 ...     ("base" = the in-tree libsso_b200.so)
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset SSO_B200_LIB; else export SSO_B200_LIB=$PWD/snark-setup-operator_b200/variants/libsso_b200_$v.so; fi
  SSO_BENCH_NOVERIFY=1 SSO_BENCH_QUICK=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$v.json"))
    ks={k["kernel"]:round(k["ms_total"]/max(1,k["launches"]),3) for k in d["roofline"]["kernels"]}
    print("$v", "ms/step %.3f" % d["ms_per_step"], "value %.3fM" % (d["value"]/1e6), "frac", d["roofline"]["frac"], ks, d["clocks"]["sm_mhz"])
except Exception as e:
    print("$v FAILED", e)
PY
done
