#!/bin/bash
# GPU tests + BASELINE config 4 at the named size     usage: tools/gpu_c4.sh <tag>
T=${1:-r02}
mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_gpu_tests.log 2>&1; tail -1 gpurun_out/${T}_gpu_tests.log
timeout 400 python tools/run_c45.py c4 --out gpurun_out/${T}_c4_n1.json 2>&1 | tail -1 | cut -c1-700
