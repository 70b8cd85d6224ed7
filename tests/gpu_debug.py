"""First-run diagnostics (not a test): prints GPU vs oracle differences stage by stage."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import obvi_b200 as ob
from oracle import oracle_lib as ol

g = ob.synth.make_graph(K=12, P=300, O=5, seed=1, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4, ltm_frac=0.5)
print(g.counts())
n = g.counts()
ref = ol.evaluate(g, apply_loss=True)
p = ob.problem_from_graph(g)
def e(a, b): return float(np.abs(a - b).max() / (1 + np.abs(b).max())) if a.size else 0.0
r, jp, jl = p.evaluate_factor_type(ob.FACTOR_REPROJECTION, n["reproj"], True); print("reproj", e(r, ref["r_reproj"]), e(jp, ref["jp_reproj"]), e(jl, ref["jl_reproj"]))
r, jo, jp = p.evaluate_factor_type(ob.FACTOR_BBOX, n["bbox"], True); print("bbox", e(r, ref["r_bbox"]), e(jo, ref["jo_bbox"]), e(jp, ref["jp_bbox"]))
r, j, _ = p.evaluate_factor_type(ob.FACTOR_SHAPE_PRIOR, n["shape"], True); print("shape", e(r, ref["r_shape"]), e(j, ref["j_shape"]))
r, j, _ = p.evaluate_factor_type(ob.FACTOR_LTM_PRIOR, n["ltm"], True); print("ltm", e(r, ref["r_ltm"]), e(j, ref["j_ltm"]))
r, j1, j2 = p.evaluate_factor_type(ob.FACTOR_REL_POSE, n["relpose"], True); print("rel", e(r, ref["r_rel"]), e(j1, ref["j1_rel"]), e(j2, ref["j2_rel"]))
c, res = p.evaluate(True); print("cost", c, ref["cost"])
for name, gg, iters in (("small", g, 12), ("C1obj", ob.synth.make_config("C1obj"), 30)):
    g2 = gg.copy()
    p = ob.problem_from_graph(gg)
    o = dict(max_num_iterations=iters, function_tolerance=1e-6, initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)
    s = p.solve(**o)
    rr = ol.solve(g2, max_num_iterations=iters, function_tolerance=1e-6, initial_radius=100.0, max_radius=1e4, use_nonmonotonic_steps=True)
    print(name, s.termination, rr["termination"], s.num_lm_steps, rr["lm_steps"], "launches", s.kernel_launches, "pcg", s.pcg_iterations_total)
    for a, b in zip(s.iterations, rr["iterations"]):
        print("%3d gpu %.10e %d pcg %4d step %.3e gmax %.3e r %.3g| cpu %.10e %d step %.3e gmax %.3e r %.3g" % (a["iteration"], a["cost"], a["successful"], a["linear_solver_iterations"], a["step_norm"], a["gradient_max_norm"], a["radius"], b["cost"], b["successful"], b["step_norm"], b["gradient_max_norm"], b["radius"]))
    print("final", s.final_cost, rr["final_cost"], "dpose", np.abs(gg.poses - g2.poses).max(), "times", s.minimizer_device_time_in_seconds, s.jacobian_evaluation_time_in_seconds, s.linear_solver_time_in_seconds, s.residual_evaluation_time_in_seconds)
