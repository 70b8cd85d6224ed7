timeout 60 python tests/gpu_time.py C1 10 2>&1 | grep -E "rep|^ *10 " | head -4
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python tests/gpu_time.py C3 50 2>&1 | grep -E "rep|^ *50 " | head -16
OBVI_PROFILE=1 timeout 120 python tests/gpu_time.py C3 50 2>&1 | grep -E "profile" | tail -${1:-12}
