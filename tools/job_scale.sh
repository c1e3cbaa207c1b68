set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
S='import sys,json; d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], d["config"]["workload"][:20], "value", round(d["value"]/1e6,2), "M ev/s ms/step", round(d["ms_per_step"],4), d["kernel_ms"], "e2e", round(d["e2e"]["ms_per_step"],4), "rate", round(d["mean_rate_hz"],1))'
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/scale_c2_n$n.json 2> gpurun_out/scale_err_$n.log
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --workload c2 --steps 100 --warmup 10 > gpurun_out/scale_c2_n$n.json 2> gpurun_out/scale_err_$n.log
    fi
    tail -1 gpurun_out/scale_c2_n$n.json | python -c "$S" || tail -5 gpurun_out/scale_err_$n.log
  fi
done
if [ $N -ge 4 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload c3 --steps 20 --warmup 5 > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_err_c3.log
  tail -1 gpurun_out/scale_c3_n$N.json | python -c "$S" || tail -5 gpurun_out/scale_err_c3.log
fi
