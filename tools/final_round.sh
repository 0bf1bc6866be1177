# Final single-GPU artefacts of the round: the driver-shaped bench line, then the profile captures of tools/profile_round.sh
mkdir -p gpurun_out/final
python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
echo "bench: rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_reference_n1.json 2> gpurun_out/final/bench_reference_n1.err
echo "reference arm: rc=$?"
bash tools/profile_round.sh
