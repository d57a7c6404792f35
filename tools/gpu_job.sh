set -x
for k in trace shade; do
  pat=k_trace; [ $k = shade ] && pat=k_shade_dense
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 60 -c 1 -o gpurun_out/r2_${k}_c4 -f python tools/profile_run.py 16 3840x2160 config=c4 > gpurun_out/r2_ncu_${k}_c4.log 2>&1
done
grep -h "STEADY\|Msamples" gpurun_out/r2_ncu_*_c4.log
python bench.py --config c4 --steps 2 --warmup 3 > gpurun_out/r2_bench_c4_a.json 2> gpurun_out/r2_bench_c4_a.err; tail -c 600 gpurun_out/r2_bench_c4_a.err; head -c 700 gpurun_out/r2_bench_c4_a.json
