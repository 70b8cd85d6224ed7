#!/bin/bash
# Round 1c: validate the split object elimination, measure it and the two experimental toggles.
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01c_tests.log 2>&1
tail -3 gpurun_out/r01c_tests.log
OBVI_PROFILE=1 timeout 240 python tests/gpu_ab.py 50 "" OBVI_OBJ_SPLIT=0 OBVI_POSE_ACCUM_SIDE=1 OBVI_DEFER_SYNC=1 OBVI_DEFER_SYNC=1,OBVI_POSE_ACCUM_SIDE=1 > gpurun_out/r01c_ab.log 2> gpurun_out/r01c_ab.err
cat gpurun_out/r01c_ab.log
grep -E "variant|pose_accum|schur_points|point_prep|zero |join\(|pcg " gpurun_out/r01c_ab.err | grep -B1 -A6 "rep 2" | head -60
( time OBVI_DEFER_SYNC=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_schedule.py -m gpu -x -q -k "solve or schedule or two_phase or pgo or tracking or points_only or zero_iter or ltm" ) > gpurun_out/r01c_tests_defer.log 2>&1
tail -3 gpurun_out/r01c_tests_defer.log
