#!/bin/bash
# Multi-GPU measurement set under torchrun (one process per GPU): contribute bench + verify_transcript workload.
#   tools/gpu_r2_multi.sh TAG N POWER CHUNKLOG
TAG=$1; N=$2; POWER=${3:-18}; CL=${4:-15}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
timeout 400 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 $RUN bench.py --gpus $N --workload verify_transcript --curve bw6_761 --power $POWER --chunk-log $CL --steps 1 --warmup 1 \
  > gpurun_out/${TAG}_vt_bw6_p${POWER}_n$N.json 2> gpurun_out/${TAG}_vt_bw6_p${POWER}_n$N.err
grep -h "Init COMPLETE" gpurun_out/${TAG}_vt_bw6_p${POWER}_n$N.err | sed 's/.*NCCL INFO //' | sort | uniq -c | head -20 > gpurun_out/${TAG}_vt_bw6_p${POWER}_n${N}_nccl_ranks.txt
python - <<PY
import json
for fn in ("gpurun_out/${TAG}_bench_n$N.json", "gpurun_out/${TAG}_vt_bw6_p${POWER}_n$N.json"):
    try: d = json.load(open(fn))
    except Exception as e: print(fn, "FAILED", e); continue
    print(fn.split("/")[-1], d["metric"], "%.5g" % d["value"], d["unit"], "n_gpus", d["n_gpus"], "e2e", (d.get("e2e") or {}).get("value"), d.get("phases_s"), d.get("nccl"))
PY
cat gpurun_out/${TAG}_vt_bw6_p${POWER}_n${N}_nccl_ranks.txt | cut -c1-200
