# ncu capture of the all-tcgen05 MLP backward (and the hybrid one beside it) on a C2 batch
ncu --set full --clock-control none --import-source on -k regex:"nerf_mlp_backward" -c 4 -f -o gpurun_out/r02e_bwd python tools/mlp_ab.py tc umma > gpurun_out/r02e_prof.log 2>&1
tail -3 gpurun_out/r02e_prof.log
