set -x
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
if [ "$N" = 1 ]; then run() { python bench.py --gpus 1 "$@"; }; fi
run --config c5 --steps 1 --warmup 3 > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err; tail -c 400 gpurun_out/r2_bench_c5_n$N.err; head -c 900 gpurun_out/r2_bench_c5_n$N.json
if [ "$N" != 1 ]; then
run --config c2 --scaling strong --steps 3 --warmup 3 > gpurun_out/r2_bench_c2strong_n$N.json 2> gpurun_out/r2_bench_c2strong_n$N.err; head -c 400 gpurun_out/r2_bench_c2strong_n$N.json
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "group_" 2>&1 | tail -3
fi
