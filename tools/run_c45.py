#!/usr/bin/env python
"""BASELINE configs 4 and 5 at their NAMED size through the CUDA backend (single GPU, or one process per GPU under torchrun).

  c4: S(2000 keyframes, 200k points, 500 objects) revealed frame by frame under the reference's schedule
      (config/base7a_2_fallback.json; include/refactoring/offline/offline_problem_runner.h:100-270,376-916;
      run_opt_utils.h:101-116): local two-phase BA on 50-frame windows, a global step (tracking solve + PGO with objects +
      re-anchoring + points-only BA) every 30 frames, PGO + two-phase visual BA at the end.  Multi-GPU: the large solves are
      sharded (obvi-slam_b200/schedule.py:ShardedGpuBackend), local windows stay on rank 0.
  c5: 4 sessions x S(1500, 150k, .) over the same 500 objects, chained through the long-term map
      (src/evaluation/ltm_trajectory_sequence_executor.py:44-83): solve session k (final two-phase global BA,
      config :64-87), extract every ellipsoid's estimate + 7x7 marginal covariance
      (long_term_object_map_extraction.cpp:362-440 -> obvi_object_covariances), hand them to session k + 1 as LTM priors
      (independent_object_map_factor.h:21-33).

  python tools/run_c45.py c4 [--frames 2000] [--out profiles/r02_c4_n1.json]
  python -m torch.distributed.run --nproc-per-node 2 ... tools/run_c45.py c4 --out profiles/r02_c4_n2.json
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import obvi_b200 as ob

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["c4", "c5"])
ap.add_argument("--frames", type=int, default=None)
ap.add_argument("--points", type=int, default=None)
ap.add_argument("--objects", type=int, default=500)
ap.add_argument("--sessions", type=int, default=4)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
S = ob.schedule
dist = template = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    template = ob.Problem(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(ob.Problem.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    template.comm_init(bytes(uid.cpu().tolist()), rank, world)


def emit(d):
    if rank == 0:
        line = json.dumps(d)
        print(line, flush=True)
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            open(a.out, "w").write(line + "\n")


def backend():
    return S.GpuBackend(ob, device=local) if world == 1 else S.ShardedGpuBackend(ob, local, dist, template)


if a.what == "c4":
    K = a.frames or 2000
    P = a.points or 100 * K                      # named size: 2000 keyframes / 200k points / 500 objects
    t = time.time()
    g = ob.synth.make_graph(K=K, P=P, O=max(1, a.objects * K // 2000), seed=a.seed, objects_on=True, relpose="all", n_const_poses=1,
                            fill=True, starved_every=10)
    gen_s = time.time() - t
    p = S.ScheduleParams()
    err = lambda: float(np.linalg.norm(g.poses[:, :3] - g.poses_gt[:, :3], axis=1).mean())
    e0 = err()
    be = backend()
    t = time.time()
    log = S.run_schedule(g, be, p)
    wall = time.time() - t
    kinds = {}
    for e in log:
        kinds[e["kind"]] = kinds.get(e["kind"], 0) + 1
    final = [e for e in log if e["kind"] == "final"][-1]
    emit(dict(config="C4: reference schedule (window 50, global step every 30 frames, two-phase BA with 10 % exclusion, PGO on global steps, PGO + visual BA at the end)",
              n_gpus=world, counts=g.counts(), generate_s=round(gen_s, 1), windows=kinds, solves=be.stats["solves"], lm_iterations=be.stats["lm_steps"],
              device_s=round(be.stats["device_s"], 3), backend_wall_s=round(be.stats["wall_s"], 2), schedule_wall_s=round(wall, 2),
              lm_iterations_per_device_s=round(be.stats["lm_steps"] / max(be.stats["device_s"], 1e-9), 1),
              structure_builds=be.stats["structure_builds"], excluded_factors=be.stats["excluded"],
              sharded_solves=be.stats.get("sharded_solves", 0), rank0_solves=be.stats.get("rank0_solves", be.stats["solves"]),
              final_ba=dict(n_reproj=final.get("n_reproj"), n_bbox=final.get("n_bbox"), costs=final["costs"]),
              mean_transl_err_before_m=round(e0, 4), mean_transl_err_after_m=round(err(), 4),
              reverted=sum(1 for e in log if "reverted" in e.get("costs", []))))
else:
    K, P = a.frames or 1500, a.points or 150000
    opts1 = S.ScheduleParams().final_phase1.as_dict(); opts2 = S.ScheduleParams().final_phase2.as_dict()
    prior = None
    sessions = []
    t_all = time.time()
    for sidx in range(a.sessions):
        g = ob.synth.make_graph(K=K, P=P, O=a.objects, seed=a.seed, noise_seed=1000 + sidx, objects_on=True, relpose="starved", n_const_poses=1,
                                fill=True, starved_every=10)
        if prior is not None:
            objs, mean, cov = prior
            g.ltm = dict(obj=objs.copy(), mean=mean.copy(), cov=cov.copy(), huber=1.0)
            g.objects[objs] = mean                      # the next session starts from the map
        be = backend()
        be.kind = "final"
        t = time.time()
        costs = be.two_phase(g, opts1, opts2, 0.1)
        solve_wall = time.time() - t
        # long-term-map extraction on rank 0 (unsharded problem at the solution): estimates + marginal covariances
        t = time.time()
        objs = np.array(sorted(set(int(o) for o in g.bbox["obj"])), np.int64)
        cov = np.zeros((len(objs), 7, 7))
        if rank == 0:
            q = ob.problem_from_graph(g, device=local)
            cov = q.object_covariances([g.objects[o] for o in objs], [g.objects[o] for o in objs])
        if world > 1:
            import torch
            tc = torch.from_numpy(np.ascontiguousarray(cov)).cuda(); dist.broadcast(tc, 0); cov = tc.cpu().numpy()
        cov_wall = time.time() - t
        prior = (objs, g.objects[objs].copy(), cov)
        eo = float(np.linalg.norm(g.objects[objs, :3] - g.objects_gt[objs, :3], axis=1).mean())
        sessions.append(dict(session=sidx, counts=g.counts(), costs=costs, lm_iterations=be.stats["lm_steps"], device_s=round(be.stats["device_s"], 3),
                             solve_wall_s=round(solve_wall, 2), ltm_objects=int(len(objs)), covariance_wall_s=round(cov_wall, 2),
                             mean_cov_trace=float(np.trace(cov, axis1=1, axis2=2).mean()), mean_object_centre_err_m=round(eo, 4),
                             mean_transl_err_m=round(float(np.linalg.norm(g.poses[:, :3] - g.poses_gt[:, :3], axis=1).mean()), 4),
                             structure_builds=be.stats["structure_builds"], excluded=be.stats["excluded"]))
        if rank == 0:
            print("session", sidx, sessions[-1], file=sys.stderr, flush=True)
    emit(dict(config="C5: sessions chained through the long-term map (estimates + 7x7 marginal covariances -> LTM prior factors of the next session); "
                     "per session the final two-phase global BA", n_gpus=world, sessions=sessions, total_wall_s=round(time.time() - t_all, 1),
              map_tightens=bool(sessions[-1]["mean_cov_trace"] < sessions[0]["mean_cov_trace"])))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
