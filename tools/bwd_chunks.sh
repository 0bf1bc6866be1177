timeout 300 python -m pytest tests/test_gpu_training.py -q -x -k "fused_table_scatter or tcgen05_wgrad" 2>&1 | tail -6
for m in 0 1; do
  echo -n "fused_scatter=$m "; NGP_B200_BWD_FUSED_SCATTER=$m python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['timing']['ms_per_step_min'], d['e2e']['ms_per_step'], d['roofline']['kernels']['nerf_mlp_backward']['ms'])"
done
