set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
S='import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["workload"][:20], "ms/step", round(d["ms_per_step"],4), d["kernel_ms"], "e2e", round(d["e2e"]["ms_per_step"],4), "rate", round(d["mean_rate_hz"],1), "frac", round(d["roofline"]["frac"],4), round(d["roofline"]["step"]["frac"],4), "deliv", d["per_step"]["deliveries"])'
for cap in 1024 768 1536; do
NC_CAND_SMEM=$cap timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
done
for cap in 1024 512 256; do
NC_CAND_SMEM=$cap timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
done
