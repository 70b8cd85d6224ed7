"""A/B timing diagnostics (not a test): one C3 graph, one Problem per environment variant (the OBVI_* switches are read when
the problem is created), 2 solves each; prints LM it/s and, with OBVI_PROFILE=1, the in-situ phase table on stderr.

  python tests/gpu_ab.py 50 "" OBVI_OBJ_SPLIT=0 OBVI_DEFER_SYNC=0,OBVI_REFACTOR_RATIO=4
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import obvi_b200 as ob

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
variants = sys.argv[2:] or [""]
g = ob.synth.make_config("C3")
x0 = (g.poses.copy(), g.points.copy(), g.objects.copy())
o = dict(max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0,
         initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)
for v in variants:
    kv = dict(x.split("=") for x in v.split(",") if x)
    os.environ.update(kv)
    p = ob.problem_from_graph(g)
    for k in kv:
        del os.environ[k]
    best = 0.0
    for rep in range(3):
        g.poses[:], g.points[:], g.objects[:] = x0
        sys.stderr.write(f"[variant {v or 'default'} rep {rep}]\n"); sys.stderr.flush()
        t = time.time(); s = p.solve(**o); w = time.time() - t
        if rep:
            best = max(best, s.num_lm_steps / s.minimizer_device_time_in_seconds)
    print(f"{v or 'default':50s} it/s {best:7.1f}  e2e {s.num_lm_steps / w:7.1f}  steps {s.num_lm_steps} ok {s.num_successful_steps} "
          f"pcg {s.pcg_iterations_total} launches {s.kernel_launches} final {s.iterations[-1]['cost']:.12e}", flush=True)
    del p
