mkdir -p gpurun_out
export YUNE_B200_LIB=$PWD/build/variants/mb4.so
for i in 1 2 3; do timeout 300 python tools/tune.py 16 '{"fused_shade":[1,1,1,1]}' 2>&1 | grep -E "fail|best"; done
unset YUNE_B200_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/tune.py 128 '{"fused_shade":[1,1]}' 2>&1 | tail -2
timeout 300 python tools/profile_run.py 64 1024 fused_shade=1 time_stages=4 | head -1
