#!/bin/bash
# row-pair Schur kernel: TMA ring depth sweep, then tests + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not termination" 2>&1 | tail -3
for st in 2 4 8; do
  echo "== OBVI_ROW_STAGES=$st"
  OBVI_ROW_STAGES=$st OBVI_PROFILE=1 timeout 120 python tests/gpu_time.py C3 50 2> gpurun_out/r02f_insitu_st$st.txt | grep -E "rep 1"
  grep profile gpurun_out/r02f_insitu_st$st.txt | grep -E "schur_points|LM steps" | tail -2
done
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02f_tests.log 2>&1
tail -5 gpurun_out/r02f_tests.log
