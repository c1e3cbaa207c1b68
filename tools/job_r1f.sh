set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S='import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["workload"][:20], "ms/step", round(d["ms_per_step"],4), d["kernel_ms"], "e2e", round(d["e2e"]["ms_per_step"],4), "rate", round(d["mean_rate_hz"],1), "frac", round(d["roofline"]["frac"],4), round(d["roofline"]["step"]["frac"],4), "deliv", d["per_step"]["deliveries"])'
timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
NC_CAND_SMEM=512 timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
NC_CAND_SMEM=512 timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_neuron_pass|k_synapse_pass" -s 1700 -c 2 -o gpurun_out/r1f_c2 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
