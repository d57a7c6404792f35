mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 150 --csv --log-file gpurun_out/launches_r1h.csv python tools/profile_run.py 32 1024 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_dense --launch-skip 30 -c 1 -o gpurun_out/prof_shade_r1h -f python tools/profile_run.py 32 1024 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_trace --launch-skip 30 -c 1 -o gpurun_out/prof_trace_r1h -f python tools/profile_run.py 32 1024 > /dev/null 2>&1
ls -la gpurun_out/*r1h*
