#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tests/gpu_time.py C3 10 2>&1 | grep -E "rep" || { echo "C3 smoke failed/hung"; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
OBVI_PROFILE=1 timeout 120 python tests/gpu_time.py C3 50 2> gpurun_out/r02j_insitu.txt | grep -E "rep 1"
grep profile gpurun_out/r02j_insitu.txt | tail -17
