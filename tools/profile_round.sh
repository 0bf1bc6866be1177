# One GPU call that produces the round's profile artefacts under gpurun_out/prof/ (copied into profiles/ afterwards):
#   launch list of 2 eager training steps + 1 density-grid update, ncu --set full of one eager step,
#   ncu metrics of the C4 hash-encoder sweep (single-pass vs level-major), ncu --set full of one inference frame's kernels
set -u
OUT=gpurun_out/prof
mkdir -p "$OUT"
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file "$OUT/launches_step.csv" python bench.py --warmup 3 --no-graph --profile 2 > "$OUT/launches_step.log" 2>&1
echo "launch list: rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o "$OUT/full_step" \
    python bench.py --warmup 3 --no-graph --profile 1 > "$OUT/full_step.log" 2>&1
echo "full step: rc=$?"
HASHENC_SWEEP_ONCE=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:hashgrid --csv --log-file "$OUT/hashenc_sweep.csv" python tools/hashenc_sweep.py > "$OUT/hashenc_sweep.log" 2>&1
echo "hashenc sweep: rc=$?"
