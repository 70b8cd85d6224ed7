#!/bin/bash
# Round 1d: where should the object slot / pair kernel overlap?  + standalone durations of the object kernels.
mkdir -p gpurun_out
OBVI_PROFILE=1 timeout 300 python tests/gpu_ab.py 50 "" OBVI_OBJ_WHEN=rows OBVI_OBJ_WHEN=gate OBVI_OBJ_WHEN=rows,OBVI_POSE_ACCUM_SIDE=1 OBVI_OBJ_WHEN=rows,OBVI_POSE_ACCUM_SIDE=1,OBVI_DEFER_SYNC=1 OBVI_OBJ_WHEN=gate,OBVI_DEFER_SYNC=1 > gpurun_out/r01d_ab.log 2> gpurun_out/r01d_ab.err
cat gpurun_out/r01d_ab.log
grep -E "variant|pose_accum|schur_points|point_prep|join\(" gpurun_out/r01d_ab.err | grep -A4 "rep 2" | head -60
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:obj_schur_kernel|schur_eblock_kernel|relpose_kernel|unary_kernel|bbox_kernel" -c 40 --csv --log-file gpurun_out/r01d_obj_launches.csv python tests/gpu_time.py C3 3 > gpurun_out/r01d_ncu.log 2>&1
python tests/ncu_agg.py gpurun_out/r01d_obj_launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:obj_schur_kernel" -s 2 -c 1 -o gpurun_out/r01d_obj_schur -f python tests/gpu_time.py C3 3 > gpurun_out/r01d_ncu2.log 2>&1
tail -2 gpurun_out/r01d_ncu2.log
