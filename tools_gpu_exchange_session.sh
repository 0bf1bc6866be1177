#!/usr/bin/env bash
# One gpurun call (>= 2 GPUs) that settles the fused gradient exchange of csrc/exchange.cu:
#   gpurun --gpus 2 --timeout 600 -- 'bash tools_gpu_exchange_session.sh 2'
# 1. parity + timing of the peer arms against the NCCL arm on the C2 buffer (tools_exchange_check.py)
# 2. the opt-in GPU test
# 3. bench.py at N ranks with each exchange, training step only
# Everything lands in gpurun_out/exchange/.  The kernel waits on its peers: a rank that never arrives makes the others trap
# after 20 s (NGP_B200_EXCHANGE_TIMEOUT_MS) with a message naming the missing rank; each step also has its own timeout.
set -u
N=${1:-2}
OUT=gpurun_out/exchange
mkdir -p "$OUT"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
for mode in peer-p2p peer; do
  timeout 120 bash -c "$(declare -f run); N=$N; run 29511 tools_exchange_check.py --mode $mode" > "$OUT/check_${mode}_n$N.log" 2>&1
  echo "check $mode: rc=$? $(grep -h '^{' "$OUT/check_${mode}_n$N.log" | tail -1)"
done
NGP_B200_TEST_EXCHANGE=1 timeout 300 python -m pytest tests/test_gpu_exchange.py -x -q > "$OUT/pytest_n$N.log" 2>&1
echo "pytest: rc=$? $(tail -1 "$OUT/pytest_n$N.log")"
# the other opt-in GPU test written without a GPU (C1 harness vs its CPU restatement); single GPU
NGP_B200_TEST_IMAGEFIT=1 timeout 300 python -m pytest tests/test_imagefit.py -x -q -m gpu > "$OUT/pytest_imagefit.log" 2>&1
echo "pytest imagefit: rc=$? $(tail -1 "$OUT/pytest_imagefit.log")"
NGP_B200_TEST_CHECKPOINT=1 timeout 300 python -m pytest tests/test_checkpoint.py -x -q -m gpu > "$OUT/pytest_checkpoint.log" 2>&1
echo "pytest checkpoint: rc=$? $(tail -1 "$OUT/pytest_checkpoint.log")"
NGP_B200_TEST_CULLING=1 timeout 300 python -m pytest tests/test_gpu_optin_ogrid.py -x -q > "$OUT/pytest_culling.log" 2>&1
echo "pytest culling: rc=$? $(tail -1 "$OUT/pytest_culling.log")"
for graph in 0 1; do  # 1: the exchange + optimizer replayed as one captured graph (NGP_B200_GRAPH_EXCHANGE)
  for mode in nccl peer; do
    NGP_B200_GRAPH_EXCHANGE=$graph timeout 300 bash -c "$(declare -f run); N=$N; run 29513 bench.py --gpus $N --steps 48 --warmup 4 --no-extras --no-cpu-baseline --exchange $mode" > "$OUT/bench_${mode}_g${graph}_n$N.log" 2>&1
    echo "bench $mode graph=$graph: rc=$? $(grep -h '^{"metric' "$OUT/bench_${mode}_g${graph}_n$N.log" | tail -1 | cut -c1-220)"
  done
done
