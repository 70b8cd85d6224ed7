#!/bin/bash
# 2-GPU confirmation: NCCL tests + bench at N = 2     usage (gpurun --gpus 2): tools/gpu_n2.sh <tag>
T=${1:-r02}
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/${T}_gpu_multi_2gpu_tests.log 2>&1
tail -2 gpurun_out/${T}_gpu_multi_2gpu_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
cut -c1-260 gpurun_out/${T}_bench_n2.json
