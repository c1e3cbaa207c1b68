set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_neuron_pass|k_synapse_pass" -s 810 -c 2 -o gpurun_out/r1h_c3 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
