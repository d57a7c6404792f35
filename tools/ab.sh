mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/tune.py 256 '{"accel":[1,1]}' 2>&1 | tail -2
timeout 300 python tools/profile_run.py 64 1024 time_stages=4 | head -1
