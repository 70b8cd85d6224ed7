python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tests/gpu_time.py C3 10 2>&1 | grep -E "rep|^ *10 " | head -16
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/l_rows.csv python tests/gpu_time.py C3 3 > /dev/null 2>&1
python tests/ncu_agg.py gpurun_out/l_rows.csv 2>/dev/null | head -${1:-12}
