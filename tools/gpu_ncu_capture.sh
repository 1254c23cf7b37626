#!/bin/bash
# One `ncu --set full` capture (with SASS/source correlation) of the dominant kernel on the bench workload.
#   tools/gpu_ncu_capture.sh TAG KERNEL_REGEX [variant]      -> gpurun_out/TAG.ncu-rep + TAG_raw.csv + TAG_source.csv
TAG=$1; KREGEX=$2; VAR=$3
mkdir -p gpurun_out
if [ -n "$VAR" ]; then export SSO_B200_LIB=$PWD/snark-setup-operator_b200/variants/libsso_b200_$VAR.so; fi
SSO_BENCH_NOVERIFY=1 SSO_BENCH_QUICK=1 SSO_BENCH_NOSAMPLER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX --launch-skip 3 --launch-count ${NCU_COUNT:-1} \
  -o gpurun_out/$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/$TAG.ncu-rep gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_source.csv
tail -3 gpurun_out/${TAG}_ncu.log
