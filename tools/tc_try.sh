timeout 300 python -m pytest tests/test_gpu_training.py -q -k "all_tcgen05" -x 2>&1 | tail -25 > gpurun_out/r02d_tc_test.log
timeout 200 python tools/mlp_ab.py tc umma > gpurun_out/r02d_mlp_ab.json 2> gpurun_out/r02d_mlp_ab.err
tail -3 gpurun_out/r02d_mlp_ab.err
