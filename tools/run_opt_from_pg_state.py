#!/usr/bin/env python
"""Counterpart of the reference's run_opt_from_pg_state (src/refactoring/run_opt_from_pg_state.cpp:67-327): load a pose-graph
state JSON, run one two-phase bundle adjustment over [min_frame, max_frame] with the reference's scope rules and solver
parameters on the CUDA backend, write the optimised state back in the same format.

  python tools/run_opt_from_pg_state.py --pg-state in.json --out out.json [--min-frame A --max-frame B] [--final]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import obvi_b200 as ob

ap = argparse.ArgumentParser()
ap.add_argument("--pg-state", required=True)
ap.add_argument("--out", required=True)
ap.add_argument("--min-frame", type=int, default=None)
ap.add_argument("--max-frame", type=int, default=None)
ap.add_argument("--final", action="store_true", help="use the final-optimisation solver parameters instead of the global-BA ones")
a = ap.parse_args()
S = ob.schedule
g, ids, extras = ob.pg_state_io.read_pose_graph_state(a.pg_state)
frames = ids["frame"]
lo = 0 if a.min_frame is None else int(np.searchsorted(frames, a.min_frame))
hi = len(frames) - 1 if a.max_frame is None else int(np.searchsorted(frames, a.max_frame, side="right")) - 1
p = S.ScheduleParams()
sub, maps = S.build_scope(g, lo, hi, p)
be = S.GpuBackend(ob)
ph = (p.final_phase1, p.final_phase2) if a.final else ((p.gba_phase1, p.gba_phase2) if S.is_global(lo, hi, p) else (p.lba_phase1, p.lba_phase2))
t = time.time()
costs = be.two_phase(sub, ph[0].as_dict(), ph[1].as_dict(), p.feature_outlier_percentage)
S.write_back(g, sub, maps)
ob.pg_state_io.write_pose_graph_state(a.out, g, ids, extras["semantic_classes"], extras["class_priors"])
print(json.dumps(dict(frames=[int(frames[lo]), int(frames[hi])], counts=sub.counts(), phase_costs=costs, excluded=be.stats["excluded"],
                      lm_steps=be.stats["lm_steps"], device_s=round(be.stats["device_s"], 4), wall_s=round(time.time() - t, 3))))
