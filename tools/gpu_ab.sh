#!/bin/bash
# A/B of variant libraries (OBVI_LIB_PATH) against the tree: in-situ phase times on C3.   usage: tools/gpu_ab.sh lib1.so lib2.so ...
mkdir -p gpurun_out
for lib in "$@" ""; do
  echo "== lib=${lib:-tree}"
  OBVI_LIB_PATH=${lib:+$PWD/$lib} OBVI_PROFILE=1 timeout 120 python tests/gpu_time.py C3 50 2> gpurun_out/ab_insitu.txt | grep -E "rep 1"
  grep profile gpurun_out/ab_insitu.txt | tail -16 | grep -E "schur_points|point_prep|pcg|LM steps"
done
