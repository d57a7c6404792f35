mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/tune.py 128 '{"fused_shade":[1,2,1,2]}' 2>&1 | tail -5
timeout 300 python tools/profile_run.py 64 1024 fused_shade=1 time_stages=4
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_dense --launch-skip 30 -c 1 -o gpurun_out/prof_shade_dense2 -f python tools/profile_run.py 32 1024 > gpurun_out/ncu_dense.log 2>&1
