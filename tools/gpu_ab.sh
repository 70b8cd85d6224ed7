timeout 60 python tests/gpu_time.py C1 10 2>&1 | grep -E "rep|^ *10 " | head -4
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python tests/gpu_time.py C3 50 2>&1 | grep -E "rep|^ *50 " | head -16
OBVI_PROFILE=1 timeout 120 python tests/gpu_time.py C3 50 2>&1 | grep -E "profile" | tail -${1:-12}
timeout 120 python -c "
import obvi_b200 as ob
g = ob.synth.make_config('C3'); p = ob.problem_from_graph(g)
sec, nb, no = p.profile_jacobian(30)
print('jacobian kernel %.1f us  %.0f GB/s algorithmic  frac(6556.2) %.3f' % (sec * 1e6, nb / sec / 1e9, nb / sec / 1e9 / 6556.2))
"
