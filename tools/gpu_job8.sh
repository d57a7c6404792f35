# 8-GPU validation of the round's final state (run under gpurun --gpus 8): group tests (2 and 4 devices), weak C2 and strong C5 / C2 lines
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "group_" 2>&1 | tail -3
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 "$@"; }
run --steps 3 --warmup 3 > gpurun_out/r2f_bench_c2_n8.json 2> gpurun_out/r2f_bench_c2_n8.err; head -c 330 gpurun_out/r2f_bench_c2_n8.json; echo
run --config c5 --steps 1 --warmup 3 > gpurun_out/r2f_bench_c5_n8.json 2> gpurun_out/r2f_bench_c5_n8.err; head -c 330 gpurun_out/r2f_bench_c5_n8.json; echo
run --config c2 --scaling strong --steps 3 --warmup 3 > gpurun_out/r2f_bench_c2strong_n8.json 2> gpurun_out/r2f_bench_c2strong_n8.err; head -c 330 gpurun_out/r2f_bench_c2strong_n8.json; echo
