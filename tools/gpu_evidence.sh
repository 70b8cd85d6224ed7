#!/bin/bash
# Round evidence: bench line, ncu launch list of the same command, ncu --set full of the top kernels.
python bench.py > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err
cut -c1-300 gpurun_out/r01b_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01b_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01b_launches_bench.log 2>&1
python tests/ncu_agg.py gpurun_out/r01b_launches_bench_steps2.csv > gpurun_out/r01b_launches_bench_steps2_summary.txt 2>/dev/null
head -12 gpurun_out/r01b_launches_bench_steps2_summary.txt
ncu --set full --clock-control none --import-source on -k "regex:reproj_jac_tma_kernel|point_prep_kernel|schur_rows_kernel|pcg_bt_resident_kernel|backsub_rows_kernel|pose_accum_tma_kernel" -s 12 -c 6 -o gpurun_out/r01b_top -f python tests/gpu_time.py C3 3 > gpurun_out/r01b_top.log 2>&1
tail -2 gpurun_out/r01b_top.log
