#!/usr/bin/env bash
# The one single-GPU call that fills in what round 2 could not measure (its GPU budget ran out before the last changes): the whole
# GPU test-suite on the final build, smoke, the default bench line, the spatial-recipe and literal-weight-law lines, the launch list
# and one `ncu --set full` capture of the three big kernels on c3.  ~15 minutes of box time.  Run under gpurun:
#   gpurun --timeout 1500 -- tools/job_first_gpu_call.sh r3a
set -u
TAG=${1:-r3a}
OUT=gpurun_out
mkdir -p $OUT
cd ${GRAFT_REPO_ROOT:-.}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $OUT/${TAG}_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 >> $OUT/${TAG}_tests.log
for WL in c3 c2 c3s c3raw; do
    timeout 600 python bench.py --workload $WL --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_c3.csv \
    python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-stdp-off --no-parity-check --min-timed-s 0 > $OUT/${TAG}_ncu_launch.log 2>&1
tools/ncu_job_r2.sh $TAG c3
tail -3 $OUT/${TAG}_tests.log
for WL in c3 c2 c3s c3raw; do head -c 400 $OUT/${TAG}_bench_${WL}.json; echo; done
