# Final-state ncu captures of the round (run under gpurun): one steady-state launch each of k_trace / k_shade_dense on C2 (16 M slots)
# and of k_trace on C4 at the bench's pool size, a launch list of a short C2 render, the device BVH builder's kernels.
set -x
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:k_trace -s 200 -c 1 -f -o gpurun_out/r2f_trace_c2 python tools/profile_run.py 1024 1024 > gpurun_out/r2f_ncu_trace_c2.log 2>&1
timeout 200 $N -k regex:k_shade_dense -s 200 -c 1 -f -o gpurun_out/r2f_shade_c2 python tools/profile_run.py 1024 1024 > gpurun_out/r2f_ncu_shade_c2.log 2>&1
timeout 300 $N -k regex:k_trace -s 40 -c 1 -f -o gpurun_out/r2f_trace_c4 python tools/profile_run.py 32 3840x2160 config=c4 pool_slots=16777216 > gpurun_out/r2f_ncu_trace_c4.log 2>&1
timeout 300 $N -k regex:k_shade_dense -s 40 -c 1 -f -o gpurun_out/r2f_shade_c4 python tools/profile_run.py 32 3840x2160 config=c4 pool_slots=16777216 > gpurun_out/r2f_ncu_shade_c4.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_c2.csv python tools/profile_run.py 64 1024 > gpurun_out/r2f_launches_c2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_tri_boxes|k_morton|k_ploc|k_level|k_emit|k_leaf|k_shade_records|k_iota|DeviceRadixSort|DeviceScan" --csv --log-file gpurun_out/r2f_bvh_build_kernels.csv python tools/profile_run.py 1 3840x2160 config=c4 > gpurun_out/r2f_bvh_build.log 2>&1
tail -3 gpurun_out/r2f_ncu_trace_c2.log gpurun_out/r2f_ncu_trace_c4.log gpurun_out/r2f_bvh_build.log
ls -la gpurun_out/r2f_*
