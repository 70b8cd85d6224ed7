#!/bin/bash
# last evidence of a round: smoke, GPU tests, bench line, small configs      usage: tools/gpu_final.sh <tag>
T=${1:-r02}
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_gpu_tests.log 2>&1; tail -1 gpurun_out/${T}_gpu_tests.log
timeout 300 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
cut -c1-200 gpurun_out/${T}_bench_n1.json
timeout 200 python tools/small_configs.py > gpurun_out/${T}_small_configs.jsonl 2> gpurun_out/${T}_small_configs.err; cut -c1-330 gpurun_out/${T}_small_configs.jsonl
