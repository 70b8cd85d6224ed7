#!/bin/bash
# round-2 call A: gpu tests, bench (new C3), in-situ profile
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a_tests.log 2>&1
tail -5 gpurun_out/r02a_tests.log
timeout 600 python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
cut -c1-600 gpurun_out/r02a_bench_n1.json
OBVI_PROFILE=1 timeout 200 python tests/gpu_time.py C3 50 2> gpurun_out/r02a_insitu_full.txt | grep -E "rep|C3" 
grep profile gpurun_out/r02a_insitu_full.txt | tail -16 > gpurun_out/r02a_insitu_profile.txt; cat gpurun_out/r02a_insitu_profile.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
