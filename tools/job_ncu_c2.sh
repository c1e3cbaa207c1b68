set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_neuron_pass -s 850 -c 1 -o gpurun_out/r1b_c2_neuron python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_n.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_synapse_pass -s 850 -c 1 -o gpurun_out/r1b_c2_synapse python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_s.log 2>&1
ls -la gpurun_out/
