set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 > gpurun_out/r1b_bench_c2_n1.log 2>&1; tail -1 gpurun_out/r1b_bench_c2_n1.log | cut -c1-2600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload c2 --steps 100 --warmup 10 > gpurun_out/r1b_bench_c2_n2.log 2>&1; tail -2 gpurun_out/r1b_bench_c2_n2.log | cut -c1-2600
timeout 600 python tools/dynamics_probe.py c3 1600 100 0.0276 2>&1 | tail -17 | cut -c1-300
