# one ncu --set full capture (with SASS-level stall samples) of the hybrid MLP backward on a C2 batch
mkdir -p gpurun_out/prof
ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_backward_umma --launch-skip 5 -c 1 -f -o gpurun_out/prof/bwd_kernel python tools/mlp_ab.py umma > gpurun_out/prof/bwd_kernel.log 2>&1
tail -2 gpurun_out/prof/bwd_kernel.log
ls -la gpurun_out/prof/
