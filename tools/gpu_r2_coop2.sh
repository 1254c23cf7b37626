#!/bin/bash
# Second A/B of round 2: unreduced ("lazy") operands in the Fq2 / Fq3 products, the cooperative MNT6 body with an affine table,
# occupancy of the cooperative BLS12-377 kernel; per-group timings; one full ncu capture of the MNT4-753 cooperative kernel.
TAG=${1:-r2n}
mkdir -p gpurun_out
V=$PWD/snark-setup-operator_b200/variants
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/${TAG}_parity.log
tail -2 gpurun_out/${TAG}_parity.log
run() {   # name lib coop curve steps
  local name=$1 lib=$2 coop=$3 curve=$4 steps=$5
  ( if [ "$lib" != base ]; then export SSO_B200_LIB=$V/libsso_b200_$lib.so; fi
    SSO_COOP_G2=$coop SSO_BENCH_NOVERIFY=1 SSO_BENCH_QUICK=1 timeout 300 python bench.py --curve $curve --steps $steps --warmup 3 --no-cpu-baseline \
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
run bls_lazy_plain   base   0 bls12_377 10
run bls_nolazy_plain nolazy 0 bls12_377 10
run bls_lazy_coop    base   1 bls12_377 10
run bls_lazy_coop_mb3 cmb3  1 bls12_377 10
run bls_lazy_plain_mb3 cmb3 0 bls12_377 10
run mnt4_coop        base   1 mnt4_753 3
run mnt4_plain       base   0 mnt4_753 3
run mnt6_coop_affine base   1 mnt6_753 3
run mnt6_coop_jac    m6noaff 1 mnt6_753 3
for c in mnt4_753 mnt6_753; do for m in 1 0; do SSO_COOP_G2=$m timeout 200 python tools/gpu_kernel_ab.py $c 15 2>&1 | tail -1 | sed -e "s/^/coop=$m /"; done; done | tee gpurun_out/${TAG}_groups.txt
# full capture of the cooperative MNT4-753 chunk kernel (chunk 2^14 keeps the replays short)
SSO_COOP_G2=1 SSO_BENCH_NOVERIFY=1 SSO_BENCH_QUICK=1 SSO_BENCH_NOSAMPLER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_exp_chunk --launch-skip 3 --launch-count 1 \
  -o gpurun_out/${TAG}_ncu_mnt4 -f python bench.py --curve mnt4_753 --chunk-log 14 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_mnt4.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_mnt4.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_mnt4_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_ncu_mnt4.ncu-rep --page source --csv > gpurun_out/${TAG}_ncu_mnt4_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_ncu_mnt4*
