# exchange check (parity + timing, both flavours) and the training-step bench at N ranks: bash tools/exch_n.sh N
N=${1:-2}
mkdir -p gpurun_out/scale
for mode in peer peer-p2p; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 tools/exchange_check.py --mode $mode > gpurun_out/scale/check2_${mode}_n$N.log 2>&1
  echo "check $mode rc=$? $(grep -h '^{' gpurun_out/scale/check2_${mode}_n$N.log | tail -1)"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus "$N" --steps 20 --warmup 5 --no-extras > gpurun_out/scale/step2_n$N.log 2>&1
echo "bench rc=$? $(grep -h '^{"metric' gpurun_out/scale/step2_n$N.log | tail -1 | cut -c1-230)"
