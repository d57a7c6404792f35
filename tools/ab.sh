mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/profile_bdpt.py 64 1024
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_bdpt --launch-skip 40 -c 1 -o gpurun_out/prof_bdpt_r1m -f python tools/profile_bdpt.py 64 1024 > /dev/null 2>&1
