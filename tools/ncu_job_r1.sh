# Round-1 profiling job (one GPU).  Tests and bench lines first (never under a profiler), then the launch list, then full captures.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r1_final_bench_c2.json 2> gpurun_out/r1_final_bench_c2.err
timeout 900 python bench.py --workload c3 --steps 50 --warmup 10 > gpurun_out/r1_final_bench_c3.json 2> gpurun_out/r1_final_bench_c3.err
timeout 900 python bench.py --workload c3 --steps 50 --warmup 10 --spinup-ms 0 --weight-scale 1 --no-cpu-baseline > gpurun_out/r1_final_bench_c3_quiet.json 2>> gpurun_out/r1_final_bench_c3.err
# launch list of the default command (c2): one pass per kernel, no replay
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1_final_launches_c2.csv python bench.py --steps 40 --warmup 5 --spinup-ms 20 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# full captures in the running regime: c2 (default) and the profiling-sized slice of c3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_neuron_pass|k_synapse_pass" -s 1700 -c 2 -o gpurun_out/r1_final_c2 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_neuron_pass|k_synapse_pass" -s 830 -c 2 -o gpurun_out/r1_final_m100 python bench.py --workload m100 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m100.log 2>&1
ls -la gpurun_out/ | tail -5
