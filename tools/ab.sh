mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for v in "" mb2; do
  if [ -n "$v" ]; then export YUNE_B200_LIB=$PWD/build/variants/$v.so; else unset YUNE_B200_LIB; fi
  echo "== variant ${v:-current}"
  timeout 300 python tools/tune.py 128 '{"accel":[1,1]}' 2>&1 | tail -2
  timeout 300 python tools/profile_run.py 64 1024 time_stages=4 | head -1
done
