#!/usr/bin/env python
"""BASELINE config 4 at scale: the reference's sliding-window two-phase schedule (obvi-slam_b200/schedule.py) on a
synthetic session through the CUDA backend.  Prints a JSON summary (not a bench.py line): windows, LM steps, device and
wall seconds, structure builds, excluded factors, trajectory error against ground truth before / after.

  python tools/run_c4.py [--frames 600] [--points-per-frame 100] [--objects 150] [--cpu-windows 0]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import obvi_b200 as ob

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=600)
ap.add_argument("--points-per-frame", type=int, default=100)
ap.add_argument("--objects", type=int, default=150)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
S = ob.schedule
t = time.time()
g = ob.synth.make_graph(K=a.frames, P=a.points_per_frame * a.frames, O=a.objects, seed=a.seed, objects_on=True, relpose="all", n_const_poses=1)
gen_s = time.time() - t
p = S.ScheduleParams()
err = lambda: float(np.linalg.norm(g.poses[:, :3] - g.poses_gt[:, :3], axis=1).mean())
e0 = err()
be = S.GpuBackend(ob)
t = time.time()
log = S.run_schedule(g, be, p)
wall = time.time() - t
kinds = {}
for e in log:
    kinds[e["kind"]] = kinds.get(e["kind"], 0) + 1
print(json.dumps(dict(config="C4 schedule (window 50, global every 30, two-phase, PGO on global steps)", counts=g.counts(), generate_s=round(gen_s, 1),
                      windows=kinds, solves=be.stats["solves"], lm_steps=be.stats["lm_steps"], device_s=round(be.stats["device_s"], 3),
                      backend_wall_s=round(be.stats["wall_s"], 2), schedule_wall_s=round(wall, 2), structure_builds=be.stats["structure_builds"],
                      excluded_factors=be.stats["excluded"], mean_transl_err_before_m=round(e0, 4), mean_transl_err_after_m=round(err(), 4),
                      reverted=sum(1 for e in log if "reverted" in e.get("costs", [])))))
