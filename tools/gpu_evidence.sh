#!/bin/bash
# Round evidence: GPU tests, bench line, ncu launch list of the same command, ncu --set full of the top kernels, in-situ phase table.
# usage: tools/gpu_evidence.sh <tag>      (files land in gpurun_out/<tag>_*)
T=${1:-r01c}
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
grep -E "passed|failed|error" gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
cut -c1-300 gpurun_out/${T}_bench_n1.json
OBVI_PROFILE=1 timeout 200 python tests/gpu_ab.py 50 "" 2> gpurun_out/${T}_insitu_profile.txt | tail -1
grep -A20 "rep 2" gpurun_out/${T}_insitu_profile.txt | grep profile > gpurun_out/${T}_insitu.tmp && mv gpurun_out/${T}_insitu.tmp gpurun_out/${T}_insitu_profile.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_launches_bench.log 2>&1
python tests/ncu_agg.py gpurun_out/${T}_launches_bench_steps2.csv > gpurun_out/${T}_launches_bench_steps2_summary.txt 2>/dev/null
head -14 gpurun_out/${T}_launches_bench_steps2_summary.txt
ncu --set full --clock-control none --import-source on -k "regex:reproj_jac_tma_kernel|point_prep_kernel|schur_rows_kernel|pcg_bt_resident_kernel|backsub_rows_kernel|pose_accum_tma_kernel|obj_schur_kernel" -s 14 -c 7 -o gpurun_out/${T}_top -f python tests/gpu_time.py C3 3 > gpurun_out/${T}_top.log 2>&1
tail -2 gpurun_out/${T}_top.log
