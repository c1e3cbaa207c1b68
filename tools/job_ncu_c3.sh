set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_neuron_pass|k_synapse_pass" -s 810 -c 2 -o gpurun_out/r1e_c3norm python bench.py --workload c3 --steps 6 --warmup 3 --spinup-ms 25 --weight-scale 0.0276 --no-cpu-baseline > gpurun_out/ncu_c3norm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_neuron_pass|k_synapse_pass" -s 12 -c 2 -o gpurun_out/r1e_c3quiet python bench.py --workload c3 --steps 6 --warmup 3 --spinup-ms 0 --no-cpu-baseline > gpurun_out/ncu_c3quiet.log 2>&1
ls -la gpurun_out/*.ncu-rep
