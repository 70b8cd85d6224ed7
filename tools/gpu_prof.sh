#!/bin/bash
# usage: tools/gpu_prof.sh <kernel-regex> <out-name> [launch-skip]   (ncu --set full of one launch per matching kernel)
ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${3:-2} -c ${4:-2} -o gpurun_out/$2 -f python tests/gpu_time.py C3 3 > gpurun_out/$2.log 2>&1
tail -3 gpurun_out/$2.log
