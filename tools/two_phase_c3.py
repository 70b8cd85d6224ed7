"""Two-phase BA at C3 scale: phase I, device top-k, in-place exclusion, phase II.  Prints timings (not a bench line)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import obvi_b200 as ob
g = ob.synth.make_config("C3")
x0 = (g.poses.copy(), g.points.copy(), g.objects.copy())
t = time.time(); p = ob.problem_from_graph(g); t_build = time.time() - t
o = dict(max_num_iterations=10, function_tolerance=1e-6, initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)
t = time.time(); s1 = p.solve(**o); t1 = time.time() - t
t = time.time(); out_rp = p.topk_outliers(ob.FACTOR_REPROJECTION, 0.1); out_bb = p.topk_outliers(ob.FACTOR_BBOX, 0.1); t_rank = time.time() - t
t = time.time()
for fid in list(out_rp) + list(out_bb):
    p.remove_residual_block(fid)
t_rm = time.time() - t
g.poses[:], g.points[:], g.objects[:] = x0
t = time.time(); s2 = p.solve(**o); t2 = time.time() - t
print(json.dumps(dict(problem_build_s=round(t_build, 3), phase1=dict(wall_s=round(t1, 3), preprocess_s=round(s1.preprocessor_time_in_seconds, 3), device_s=round(s1.minimizer_device_time_in_seconds, 4), final_cost=s1.final_cost),
                      rank_s=round(t_rank, 3), excluded=[len(out_rp), len(out_bb)], remove_calls_s=round(t_rm, 3),
                      phase2=dict(wall_s=round(t2, 3), preprocess_s=round(s2.preprocessor_time_in_seconds, 3), device_s=round(s2.minimizer_device_time_in_seconds, 4), initial_cost=s2.initial_cost, final_cost=s2.final_cost),
                      structure_builds=p.num_structure_builds())))
