"""Jacobian-kernel timing (not a test): 30 back-to-back launches on C3 through obvi_profile_jacobian."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import obvi_b200 as ob
g = ob.synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C3"); p = ob.problem_from_graph(g)
sec, nb, no = p.profile_jacobian(30)
print('jacobian kernel %.1f us  %.0f GB/s algorithmic  frac(6533.5) %.3f  obs %d' % (sec * 1e6, nb / sec / 1e9, nb / sec / 1e9 / 6533.5, no))
