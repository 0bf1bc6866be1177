# 8-GPU check of the gradient exchange: the fused NVLink kernel (default) against the NCCL arm, training step only,
# then the default bench line (with the tile-sharded render) as the driver will run it.
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
mkdir -p gpurun_out/scale
for mode in peer nccl; do
  timeout 240 bash -c "$(declare -f run); N=$N; run 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline --exchange $mode" > gpurun_out/scale/bench_${mode}_n$N.log 2>&1
  echo "bench $mode: rc=$? $(grep -h '^{"metric' gpurun_out/scale/bench_${mode}_n$N.log | tail -1 | cut -c1-260)"
done
timeout 120 bash -c "$(declare -f run); N=$N; run 29515 tools/exchange_check.py --mode peer" > gpurun_out/scale/check_peer_n$N.log 2>&1
echo "check peer: rc=$? $(grep -h '^{' gpurun_out/scale/check_peer_n$N.log | tail -1)"
timeout 400 bash -c "$(declare -f run); N=$N; run 29517 bench.py --gpus $N --steps 20 --warmup 5" > gpurun_out/scale/bench_default_n$N.log 2>&1
echo "bench default: rc=$? $(grep -h '^{"metric' gpurun_out/scale/bench_default_n$N.log | tail -1 | cut -c1-200)"
