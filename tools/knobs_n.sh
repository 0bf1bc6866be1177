# the training-step bench at N ranks under a few knob settings: bash tools/knobs_n.sh N
N=${1:-8}
mkdir -p gpurun_out/scale
run() {
  tag=$1; shift
  env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus "$N" --steps 20 --warmup 5 --no-extras > gpurun_out/scale/knob_${tag}_n$N.log 2>&1
  echo "$tag rc=$? $(grep -h '^{"metric' gpurun_out/scale/knob_${tag}_n$N.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["timing"]["ms_per_step_min"], d["e2e"]["ms_per_step"])')"
}
run march2 NGP_B200_MARCH_CTAS_PER_SM=2
