# the single-GPU training step under one runtime setting per group of arguments:
#   bash tools/knobs_1.sh tag VAR=value [VAR=value ...] [-- tag VAR=value ...]
# e.g. bash tools/knobs_1.sh march2 NGP_B200_MARCH_CTAS_PER_SM=2 -- c2 NGP_B200_BWD_FUSED_SCATTER=0 NGP_B200_BWD_CHUNKS=2
mkdir -p gpurun_out/knobs
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --no-ref-gpu > gpurun_out/knobs/$tag.log 2>&1
  echo "$tag rc=$? $(grep -h '^{"metric' gpurun_out/knobs/$tag.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["timing"]["ms_per_step_min"], d["e2e"]["ms_per_step"])')"
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
