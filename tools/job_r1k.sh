set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
S='import sys,json; d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], d["config"]["workload"][:20], "ms/step", round(d["ms_per_step"],4), d["kernel_ms"], "e2e", round(d["e2e"]["ms_per_step"],4), "rate", round(d["mean_rate_hz"],1), "deliv", d["per_step"]["deliveries"])'
timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
NC_FORCE_COARSE_MASK=1 timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
NC_FORCE_COARSE_MASK=1 timeout 600 python bench.py --workload m100 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
timeout 600 python bench.py --workload m100 --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload c2 --steps 100 --warmup 10 2>&1 | tail -1 | python -c "$S"
