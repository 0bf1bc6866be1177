# the driver-shaped bench line at N ranks of one box: bash tools/scale_n.sh N
N=${1:-2}
mkdir -p gpurun_out/scale
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus "$N" --steps 20 --warmup 5 \
  > gpurun_out/scale/final_n$N.log 2> gpurun_out/scale/final_n$N.err
echo "rc=$? $(grep -h '^{"metric' gpurun_out/scale/final_n$N.log | tail -1 | cut -c1-200)"
