for w in "7,7" "8,6" "9,5" "10,4" "7,5,2" "6,5,3" "5,4,3,2"; do
  echo -n "waves=$w "; NGP_B200_BWD_WAVES=$w python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['timing']['ms_per_step_min'], d['e2e']['ms_per_step'])"
done
