# prefetched-march grid size (persistent CTAs per SM) at N = 1: two more settings, then the driver-shaped line with the best one
mkdir -p gpurun_out/knobs gpurun_out/final
best=2; best_ms=0.4438
for c in 3 4; do
  NGP_B200_MARCH_CTAS_PER_SM=$c timeout 150 python bench.py --steps 20 --warmup 5 --no-extras --no-ref-gpu > gpurun_out/knobs/march$c.log 2>&1
  ms=$(grep -h '^{"metric' gpurun_out/knobs/march$c.log | tail -1 | python -c 'import sys,json; print(json.loads(sys.stdin.read())["ms_per_step"])')
  echo "march$c $ms"
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$best_ms') - 0.002 else 1)"; then best=$c; best_ms=$ms; fi
done
echo "best=$best ($best_ms)"
NGP_B200_MARCH_CTAS_PER_SM=$best timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_n1_march$best.json 2> gpurun_out/final/bench_n1_march.err
echo "final rc=$? $(grep -h '^{"metric' gpurun_out/final/bench_n1_march$best.json | tail -1 | cut -c1-220)"
