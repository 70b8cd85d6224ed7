#!/bin/bash
mkdir -p gpurun_out
for cfg in "build_ab/lib_minb8.so 2" "build_ab/lib_minb8.so 4" " 2"; do
  set -- $cfg; lib=$1; st=$2; if [ -z "$st" ]; then st=$lib; lib=""; fi
  echo "== lib=${lib:-tree} OBVI_ROW_STAGES=$st"
  OBVI_LIB_PATH=${lib:+$PWD/$lib} OBVI_ROW_STAGES=$st OBVI_PROFILE=1 timeout 120 python tests/gpu_time.py C3 50 2> gpurun_out/r02h_insitu.txt | grep -E "rep 1"
  grep profile gpurun_out/r02h_insitu.txt | grep -E "schur_points|LM steps" | tail -2
done
