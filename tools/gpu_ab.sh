python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tests/gpu_time.py C3 50 2>&1 | grep -E "rep" | head -16
OBVI_PROFILE=1 python tests/gpu_time.py C3 50 2>&1 | grep -E "profile" | tail -14
