#!/usr/bin/env bash
# Weak-scaling runs of round 2 on N GPUs of one box: bench.py under torchrun for the given workloads.
# usage: tools/job_scale_r2.sh <tag> <N> <workload> [<workload> ...]
set -u
TAG=$1; N=$2; shift 2
PORT=29520
for WL in "$@"; do
    PORT=$((PORT + 1))
    if [ "$N" = "1" ]; then
        timeout 900 python bench.py --gpus 1 --workload $WL --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_${WL}_n${N}.json 2> gpurun_out/${TAG}_${WL}_n${N}.err
    else
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --workload $WL --steps 50 --warmup 5 > gpurun_out/${TAG}_${WL}_n${N}.json 2> gpurun_out/${TAG}_${WL}_n${N}.err
    fi
    echo "== $WL n=$N rc=$?"; tail -c 600 gpurun_out/${TAG}_${WL}_n${N}.json | head -c 300; echo; tail -2 gpurun_out/${TAG}_${WL}_n${N}.err
done
