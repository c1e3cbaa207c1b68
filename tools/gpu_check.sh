# One-GPU check of a build: parity tests, smoke, default bench line and the C3 line (run under gpurun).
set -x
cd ${GRAFT_REPO_ROOT:-.}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1200
timeout 900 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1200
