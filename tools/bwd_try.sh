# MLP backward experiment loop: parity tests of every arm, isolated timing, the step
mkdir -p gpurun_out/bwd
timeout 600 python -m pytest tests -x -q -m gpu -k "mlp or backward or train_step or fused" > gpurun_out/bwd/tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/bwd/tests.log)"
timeout 300 python tools/mlp_ab.py umma > gpurun_out/bwd/ab.log 2>&1
echo "ab rc=$? $(tail -1 gpurun_out/bwd/ab.log)"
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-ref-gpu > gpurun_out/bwd/bench.log 2>&1
echo "bench rc=$? $(grep -h '^{"metric' gpurun_out/bwd/bench.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["timing"]["ms_per_step_min"], d["roofline"]["kernel"], d["roofline"]["frac"])')"
