"""ncu target: a few launches of the Jacobian-evaluation kernel on the C3 graph (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import obvi_b200 as ob
g = ob.synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C3")
p = ob.problem_from_graph(g)
print(p.profile_jacobian(reps=3))
