#!/bin/bash
# quick confirmation after a kernel change: GPU tests + in-situ phase table on C3
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
OBVI_PROFILE=1 timeout 120 python tests/gpu_time.py C3 50 2> gpurun_out/quick_insitu.txt | grep -E "rep 1"
grep profile gpurun_out/quick_insitu.txt | tail -16
