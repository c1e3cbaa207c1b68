set -x
cd $GRAFT_REPO_ROOT
S='import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["workload"][:20], "ms/step", round(d["ms_per_step"],4), d["kernel_ms"], "e2e", round(d["e2e"]["ms_per_step"],4), "rate", round(d["mean_rate_hz"],1), "frac", round(d["roofline"]["frac"],4), round(d["roofline"]["step"]["frac"],4), "deliv", d["per_step"]["deliveries"])'
timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --spinup-ms 0 --weight-scale 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
