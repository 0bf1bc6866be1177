# the training-step bench at N ranks under one knob setting per line of tools/knobs_n.txt-style arguments:
#   bash tools/knobs_n.sh N tag VAR=value [VAR=value ...] [-- tag VAR=value ...]
# e.g. bash tools/knobs_n.sh 8 b64c1 NGP_B200_EXCHANGE_BLOCKS=64 NGP_B200_BWD_CHUNKS=1 -- fused NGP_B200_BWD_FUSED_SCATTER=1
# (the sweeps recorded in profiles/exchange_knobs_r02.txt and profiles/backward_knobs_r02.txt were taken this way)
N=${1:-8}; shift
mkdir -p gpurun_out/scale
run() {
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus "$N" --steps 20 --warmup 5 --no-extras > gpurun_out/scale/knob_${tag}_n$N.log 2>&1
  echo "$tag rc=$? $(grep -h '^{"metric' gpurun_out/scale/knob_${tag}_n$N.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["timing"]["ms_per_step_min"], d["e2e"]["ms_per_step"])')"
}
args=()
for a in "$@" --; do
  if [ "$a" = "--" ]; then
    [ ${#args[@]} -gt 0 ] && run "${args[@]}"
    args=()
  else
    args+=("$a")
  fi
done
