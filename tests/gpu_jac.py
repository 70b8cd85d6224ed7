"""Timing of the Jacobian-evaluation kernel revisions (not a test)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import obvi_b200 as ob
    g = ob.synth.make_config("C3")
    p = ob.problem_from_graph(g)
    sec, nb, no = p.profile_jacobian(reps=50)
    c, _ = p.evaluate(apply_loss_function=True, residuals=False)
    print(json.dumps(dict(mode=os.environ.get("OBVI_JAC", "persistent"), us=sec * 1e6, gbs=nb / sec / 1e9, frac=nb / sec / 1e9 / 6556.2, cost=c)))
else:
    for m, rot in (("tma", "0"), ("tma2", "0")):
        print(rot, subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, OBVI_JAC=m, OBVI_JAC_ROT=rot), capture_output=True, text=True).stdout.strip())
