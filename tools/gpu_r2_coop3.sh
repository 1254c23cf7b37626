#!/bin/bash
# Third A/B of round 2: unreduced operands in the point formulas (SSO_LAZY_EC) on the prime fields and the cooperative towers;
# variants: dedicated 24-limb squaring in the point formulas, 168-register cap on MNT4-753, squarings traded on 12 limbs.
TAG=${1:-r2p}
mkdir -p gpurun_out
V=$PWD/snark-setup-operator_b200/variants
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/${TAG}_parity.log
tail -2 gpurun_out/${TAG}_parity.log
run() {   # name lib coop curve steps
  local name=$1 lib=$2 coop=$3 curve=$4 steps=$5
  ( if [ "$lib" != base ]; then export SSO_B200_LIB=$V/libsso_b200_$lib.so; fi
    if [ "$coop" != default ]; then export SSO_COOP_G2=$coop; fi
    SSO_BENCH_NOVERIFY=1 SSO_BENCH_QUICK=1 timeout 300 python bench.py --curve $curve --steps $steps --warmup 3 --no-cpu-baseline \
      > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err )
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
run bls_lazyec       base     default bls12_377 10
run bls_lazyec_trade tradeall default bls12_377 10
run bw6_lazyec       base     default bw6_761 5
run mnt4_lazyec      base     default mnt4_753 3
run mnt6_lazyec      base     default mnt6_753 3
run mnt4_old_sqr24   sqr24    1 mnt4_753 3
run mnt4_old_mb3     m4mb3    1 mnt4_753 3
