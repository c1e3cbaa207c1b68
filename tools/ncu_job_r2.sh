#!/usr/bin/env bash
# ncu captures of round 2 (one GPU): full sets of the staging kernel, the neuron pass and the synapse kernels on the
# profiling-sized slice of C3 (m100) and on C3 itself, plus the launch list of the default bench command.
# usage: tools/ncu_job_r2.sh <tag> [workload=m100]
set -u
TAG=${1:-r2}
WL=${2:-m100}
OUT=gpurun_out
SKIP=$(( (400 + 3) * 2 ))   # 25 ms spin-up = 400 steps, 3 warm-up steps; two matching launches per step for the -k filter below
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_stage|k_neuron_pass' -s $SKIP -c 4 -o $OUT/${TAG}_${WL}_neuron -f \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline --no-stdp-off --no-parity-check --min-timed-s 0 > $OUT/${TAG}_${WL}_ncu_neuron.log 2>&1
SKIP3=$(( (400 + 3) * 3 ))
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_syn_' -s $SKIP3 -c 6 -o $OUT/${TAG}_${WL}_syn -f \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline --no-stdp-off --no-parity-check --min-timed-s 0 > $OUT/${TAG}_${WL}_ncu_syn.log 2>&1
