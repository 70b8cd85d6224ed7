#!/bin/bash
# LL (flag-in-data) PCG: smoke with a short timeout first, then parity tests, then A/B against the counter/flag kernel
mkdir -p gpurun_out
timeout 90 python tests/gpu_time.py C1 5 2>&1 | tail -3 || { echo "C1 smoke failed/hung"; exit 1; }
timeout 120 python tests/gpu_time.py C3 10 2>&1 | grep -E "rep" || { echo "C3 smoke failed/hung"; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for mode in flags ll; do
  echo "== OBVI_PCG=$mode"
  OBVI_PCG=$mode OBVI_PROFILE=1 timeout 120 python tests/gpu_time.py C3 50 2> gpurun_out/r02i_insitu_$mode.txt | grep -E "rep 1"
  grep profile gpurun_out/r02i_insitu_$mode.txt | grep -E "pcg|LM steps" | tail -2
done
