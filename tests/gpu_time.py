"""Timing diagnostics (not a test): per-phase device times of a solve on a named config."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import obvi_b200 as ob

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-12
t = time.time(); g = ob.synth.make_config(name); print(name, g.counts(), "gen %.1fs" % (time.time() - t))
t = time.time(); p = ob.problem_from_graph(g); print("problem build %.2fs" % (time.time() - t))
o = dict(max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0,
         initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1, pcg_relative_tolerance=tol)
x0 = (g.poses.copy(), g.points.copy(), g.objects.copy())
for rep in range(2):
    g.poses[:], g.points[:], g.objects[:] = x0
    t = time.time(); s = p.solve(**o); w = time.time() - t
    print("rep", rep, "wall %.3f" % w, "prep %.3f" % s.preprocessor_time_in_seconds, "device loop %.4f" % s.minimizer_device_time_in_seconds,
          "jac %.4f lin %.4f res %.4f" % (s.jacobian_evaluation_time_in_seconds, s.linear_solver_time_in_seconds, s.residual_evaluation_time_in_seconds),
          "lm steps", s.num_lm_steps, "ok", s.num_successful_steps, "pcg its", s.pcg_iterations_total, "launches", s.kernel_launches,
          "it/s %.1f" % (s.num_lm_steps / s.minimizer_device_time_in_seconds))
for a in s.iterations: print("%3d %.10e ok=%d pcg=%d step %.3e r=%.3g" % (a["iteration"], a["cost"], a["successful"], a["linear_solver_iterations"], a["step_norm"], a["radius"]))
