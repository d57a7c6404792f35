mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',d['e2e']); print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','traffic','avg_launch_ms','trace_share_of_step')}); print('shade',d['roofline_shade']); print('cpu',d['cpu_baseline']); print(d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-600
