# frame breakdown (torch profiler) of both renderers + one ncu --set full capture of the persistent frame kernel
mkdir -p gpurun_out/prof
python tools/profile_render.py 300 > gpurun_out/prof/render_frame_persistent.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:nerf_render_frame -c 1 -f -o gpurun_out/prof/frame_kernel python tools/profile_render.py 300 > gpurun_out/prof/frame_kernel.log 2>&1
tail -2 gpurun_out/prof/frame_kernel.log
