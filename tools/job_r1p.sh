set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
S='import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["workload"][:20], "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), d["kernel_ms"])'
timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
NC_NEURON_VARIANT=dense timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --spinup-ms 0 --weight-scale 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
NC_NEURON_VARIANT=sparse timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$S"
