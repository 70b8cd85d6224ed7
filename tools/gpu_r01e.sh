#!/bin/bash
# Round 1e: 256-thread / 64-register object slot-pair kernel; candidates for the default switches.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01e_tests.log 2>&1
tail -4 gpurun_out/r01e_tests.log | head -2
( time OBVI_DEFER_SYNC=1 OBVI_POSE_ACCUM_SIDE=1 timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01e_tests_defer_side.log 2>&1
tail -4 gpurun_out/r01e_tests_defer_side.log | head -2
OBVI_PROFILE=1 timeout 300 python tests/gpu_ab.py 50 "" OBVI_DEFER_SYNC=1,OBVI_POSE_ACCUM_SIDE=1 OBVI_DEFER_SYNC=1 OBVI_OBJ_WHEN=rows,OBVI_DEFER_SYNC=1,OBVI_POSE_ACCUM_SIDE=1 "" OBVI_DEFER_SYNC=1,OBVI_POSE_ACCUM_SIDE=1 > gpurun_out/r01e_ab.log 2> gpurun_out/r01e_ab.err
cat gpurun_out/r01e_ab.log
grep -E "variant|pose_accum|schur_points|point_prep|join\(" gpurun_out/r01e_ab.err | grep -A4 "rep 2" | head -40
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:obj_schur_kernel|schur_eblock_kernel" -c 12 --csv --log-file gpurun_out/r01e_obj_launches.csv python tests/gpu_time.py C3 3 > gpurun_out/r01e_ncu.log 2>&1
python tests/ncu_agg.py gpurun_out/r01e_obj_launches.csv
