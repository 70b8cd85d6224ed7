#!/bin/bash
# Round 1f: REDG instead of ATOMG in the object kernels, unrolled single-division SPD inverse, deferred read-back on by default.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01f_tests.log 2>&1
grep -E "passed|failed|error" gpurun_out/r01f_tests.log
OBVI_PROFILE=1 timeout 300 python tests/gpu_ab.py 50 "" OBVI_OBJ_SPLIT=0,OBVI_DEFER_SYNC=0 "" OBVI_DEFER_SYNC=0 > gpurun_out/r01f_ab.log 2> gpurun_out/r01f_ab.err
cat gpurun_out/r01f_ab.log
grep -E "variant|pose_accum|schur_points|point_prep|zero |finish|factor_bt|pcg |backsub|candidate|jacobian|pose_cam" gpurun_out/r01f_ab.err | grep -A12 "default rep 2" | head -30
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:obj_schur_kernel|schur_eblock_kernel|minv_kernel|bt_invert8" -c 16 --csv --log-file gpurun_out/r01f_obj_launches.csv python tests/gpu_time.py C3 3 > gpurun_out/r01f_ncu.log 2>&1
python tests/ncu_agg.py gpurun_out/r01f_obj_launches.csv
