#!/bin/bash
mkdir -p gpurun_out
OBVI_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/n2prof.json 2> gpurun_out/n2prof.err
grep "obvi profile" gpurun_out/n2prof.err | tail -60 > gpurun_out/r02f_insitu_profile_n2.txt
tail -36 gpurun_out/r02f_insitu_profile_n2.txt
