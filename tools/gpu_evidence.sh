#!/bin/bash
# Round evidence: GPU tests, bench line, in-situ phase table, ncu launch list of the bench command, ncu --set full of the Jacobian
# kernel, schedule-prefix parity, small configs.   usage: tools/gpu_evidence.sh <tag>      (files land in gpurun_out/<tag>_*)
T=${1:-r02}
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests.log | head -2
timeout 300 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
cut -c1-200 gpurun_out/${T}_bench_n1.json
OBVI_PROFILE=1 timeout 100 python tests/gpu_time.py C3 50 2> gpurun_out/${T}_insitu_full.txt | grep -E "rep 1"
grep profile gpurun_out/${T}_insitu_full.txt | tail -16 > gpurun_out/${T}_insitu_profile.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_launches_bench.log 2>&1
python tests/ncu_agg.py gpurun_out/${T}_launches_bench_steps2.csv > gpurun_out/${T}_launches_bench_steps2_summary.txt 2>/dev/null
head -8 gpurun_out/${T}_launches_bench_steps2_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:reproj_jac_fused_kernel|point_prep_kernel|schur_rows_kernel|backsub_rows_kernel|pcg_bt_ll_kernel" -s 10 -c 6 -o gpurun_out/${T}_top -f python tests/gpu_time.py C3 3 > gpurun_out/${T}_top.log 2>&1
tail -1 gpurun_out/${T}_top.log
timeout 200 python tools/run_c4_prefix_parity.py --frames 60 --tight --oracle-in tests/golden/c4_prefix60_tight_oracle.json --out gpurun_out/${T}_c4_prefix60_tight_parity.json 2>&1 | tail -1 | cut -c1-300
timeout 200 python tools/small_configs.py > gpurun_out/${T}_small_configs.jsonl 2> gpurun_out/${T}_small_configs.err; cat gpurun_out/${T}_small_configs.jsonl | cut -c1-400
if [ "$2" = "c4" ]; then timeout 400 python tools/run_c45.py c4 --out gpurun_out/${T}_c4_n1.json 2>&1 | tail -1 | cut -c1-600; fi
