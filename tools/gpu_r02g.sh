#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:schur_rows -s 2 -c 1 -f -o gpurun_out/r02g_schur_rows_tma python tests/gpu_time.py C3 3 > gpurun_out/r02g_ncu.log 2>&1
tail -3 gpurun_out/r02g_ncu.log
