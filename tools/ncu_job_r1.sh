set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# plain bench lines first (never under a profiler)
timeout 900 python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/r1_bench_c3.log 2>&1
tail -1 gpurun_out/r1_bench_c3.log | cut -c1-2500
timeout 600 python bench.py --workload m100 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r1_bench_m100.log 2>&1
tail -1 gpurun_out/r1_bench_m100.log | cut -c1-1500
# launch list (one pass per kernel, no replay): m100
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_m100.csv python bench.py --workload m100 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m100_bench.log 2>&1
# full sets on the profiling-sized slice (100M synapses, state 2.8 GB > L2)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_neuron_pass -s 8 -c 2 -o gpurun_out/r1_prof_neuron python bench.py --workload m100 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m100_n.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_synapse_pass -s 8 -c 2 -o gpurun_out/r1_prof_synapse python bench.py --workload m100 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m100_s.log 2>&1
ls -la gpurun_out/
