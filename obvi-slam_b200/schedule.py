"""Host-side mirror of the reference's optimisation SCHEDULE over this package's backend.

What the reference does around ceres::Solve (all host logic, restated here so that BASELINE config 4 -- the sliding-window,
two-phase schedule -- can be run and parity-checked end to end):

  * window rule                      include/run_optimization_utils/run_opt_utils.h:101-116, optimization_runner.h:195-203
  * frame addition (pose chained from the previous optimised pose)   pose_graph_frame_data_adder.h:199-207
  * scope -> factor set              include/refactoring/optimization/object_pose_graph_optimizer.h:196-405
      (features with >= min observations inside the window, rel-pose factors only for feature-starved frames,
       objects with >= min observations or a long-term-map prior, their shape / LTM priors)
  * constant blocks                  object_pose_graph_optimizer.h:440-472
  * two-phase BA with outlier exclusion and jump reversion           include/refactoring/offline/offline_problem_runner.h:376-916
  * "global" step: tracking solve + PGO-with-objects + feature re-anchoring + points-only BA
                                     offline_problem_runner.h:438-520, pose_graph_plus_objects_optimizer.h:24-353

The numerical work of every solve is done by `backend` (GpuBackend: the CUDA path through the C ABI, with the outlier
ranking on the device and the exclusion done in place).  A backend is any object with solve(sub, opts) and
two_phase(sub, opts1, opts2, fraction); tests/helpers.py holds one that runs the CPU oracle to check the composition.
Graphs are synth.FactorGraph objects; a window is optimised on an extracted sub-graph whose arrays are the parameter
blocks, and the results are copied back -- the analogue of the reference's pose-graph nodes being updated in place.
"""
from __future__ import annotations

import dataclasses
import time

import numpy as np

from . import synth


@dataclasses.dataclass
class SolverParams:
    max_num_iterations: int = 50
    function_tolerance: float = 1e-3
    gradient_tolerance: float = 1e-10
    parameter_tolerance: float = 1e-8
    initial_trust_region_radius: float = 100.0
    max_trust_region_radius: float = 1e4
    use_nonmonotonic_steps: int = 1

    def as_dict(self):
        return dataclasses.asdict(self)


@dataclasses.dataclass
class ScheduleParams:
    """Defaults = config/base7a_2_fallback.json."""
    local_ba_window_size: int = 50
    global_ba_frequency: int = 30
    poses_prior_to_window_to_keep_constant: int = 5
    min_low_level_feature_observations: int = 5
    min_low_level_feature_observations_per_frame: int = 50
    min_object_observations: int = 10
    feature_outlier_percentage: float = 0.1
    two_phase: bool = True
    allow_reversion_after_detecting_jumps: bool = True
    consecutive_pose_transl_tol: float = 1.0
    consecutive_pose_orient_tol: float = np.pi
    use_pose_graph_on_global_ba: bool = True          # mid-run global steps: PGO instead of visual BA
    use_visual_features_on_global_ba: bool = False
    use_pose_graph_on_final_global_ba: bool = True    # last frame: PGO, then visual BA
    use_visual_features_on_final_global_ba: bool = True
    pgo_huber: float = 5.0
    pgo_cov_multiplier: float = 0.1
    enable_visual_non_opt_feature_adjustment_post_pgo: bool = True
    enable_visual_feats_only_opt_post_pgo: bool = True
    lba_phase1: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(50, 1e-3))
    lba_phase2: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(100, 1e-4))
    gba_phase1: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(250, 1e-6))
    gba_phase2: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(250, 1e-6))
    final_phase1: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(300, 1e-6))
    final_phase2: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(300, 1e-6))
    pgo: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(250, 1e-6))
    final_pgo: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(300, 1e-6))
    pre_pgo_tracking: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(50, 1e-3))
    post_pgo_vf_adjustment: SolverParams = dataclasses.field(default_factory=lambda: SolverParams(50, 1e-3))


def window_start(next_frame, max_frame, p: ScheduleParams):
    """run_opt_utils.h:101-116."""
    if next_frame == max_frame or next_frame % p.global_ba_frequency == 0 or next_frame < p.local_ba_window_size:
        return 0
    return next_frame - p.local_ba_window_size


def is_global(start, next_frame, p: ScheduleParams):
    """optimization_runner.h:195-203."""
    return next_frame - start > p.local_ba_window_size


# ----------------------------------------------------------------------------- pose algebra on (t, axis-angle) rows
def _Rt(poses):
    return synth.rotvec_to_mat(poses[..., 3:6]), poses[..., 0:3]


def compose(a, b):
    """T_a * T_b for pose rows."""
    Ra, ta = _Rt(a); Rb, tb = _Rt(b)
    return np.concatenate([np.einsum("...ij,...j->...i", Ra, tb) + ta, synth.mat_to_rotvec(np.einsum("...ij,...jk->...ik", Ra, Rb))], axis=-1)


def relative(a, b):
    """T_a^-1 * T_b."""
    Ra, ta = _Rt(a); Rb, tb = _Rt(b)
    return np.concatenate([np.einsum("...ji,...j->...i", Ra, tb - ta), synth.mat_to_rotvec(np.einsum("...ji,...jk->...ik", Ra, Rb))], axis=-1)


# ----------------------------------------------------------------------------- scope -> sub-graph
def build_scope(g, start, nxt, p: ScheduleParams, *, include_visual=True, include_objects=True, fix_poses=False,
                fix_objects=False, n_const=None, relpose_mode="starved", first_obs=None):
    """object_pose_graph_optimizer.h:196-405 + :440-472 on array-backed graphs.  Returns (sub graph, index maps)."""
    rp, bb, sh, lt, rl = g.reproj, g.bbox, g.shape, g.ltm, g.relpose
    in_win = lambda f: (f >= start) & (f <= nxt)
    # visual factors in the window, features with enough observations inside it
    if include_visual and len(rp["pose"]):
        if getattr(g, "_rp_sorted", None) is None:          # observations in pose order: a window is one slice
            g._rp_sorted = bool(np.all(np.diff(rp["pose"]) >= 0))
        if g._rp_sorted:
            lo, hi = np.searchsorted(rp["pose"], [start, nxt + 1])
            cnt = np.bincount(rp["point"][lo:hi], minlength=len(g.points))
            rp_idx = lo + np.nonzero(cnt[rp["point"][lo:hi]] >= p.min_low_level_feature_observations)[0]
        else:
            m = in_win(rp["pose"])
            cnt = np.bincount(rp["point"][m], minlength=len(g.points))
            m &= cnt[rp["point"]] >= p.min_low_level_feature_observations
            rp_idx = np.nonzero(m)[0]
    else:
        rp_idx = np.zeros(0, np.int64)
    # rel-pose factors: only for window frames with too few feature observations (both ends inside the window)
    if relpose_mode == "starved":
        per_frame = np.bincount(rp["pose"][rp_idx], minlength=len(g.poses)) if len(rp_idx) else np.zeros(len(g.poses), np.int64)
        starved = per_frame < p.min_low_level_feature_observations_per_frame
        rl_m = in_win(rl["p1"]) & in_win(rl["p2"]) & (starved[rl["p1"]] | starved[rl["p2"]]) if len(rl["p1"]) else np.zeros(0, bool)
    else:
        rl_m = np.zeros(len(rl["p1"]), bool)
    rl_idx = np.nonzero(rl_m)[0]
    # objects: enough bbox observations in the window, or a long-term-map prior
    if include_objects and len(bb["obj"]):
        bm = in_win(bb["pose"])
        ocnt = np.bincount(bb["obj"][bm], minlength=len(g.objects))
        keep_obj = ocnt >= p.min_object_observations
        if len(lt["obj"]):
            keep_obj[lt["obj"]] = True
        bm &= keep_obj[bb["obj"]]
    else:
        bm = np.zeros(len(bb["obj"]), bool)
        keep_obj = np.zeros(len(g.objects), bool)
        if include_objects and len(lt["obj"]):
            keep_obj[lt["obj"]] = True
    bb_idx = np.nonzero(bm)[0]
    sh_idx = np.nonzero(keep_obj[sh["obj"]])[0] if include_objects and len(sh["obj"]) else np.zeros(0, np.int64)
    lt_idx = np.nonzero(keep_obj[lt["obj"]])[0] if include_objects and len(lt["obj"]) else np.zeros(0, np.int64)

    pose_ids = np.arange(start, nxt + 1)
    point_ids = np.unique(rp["point"][rp_idx]) if len(rp_idx) else np.zeros(0, np.int64)
    obj_ids = np.nonzero(keep_obj)[0]
    pmap = -np.ones(len(g.poses), np.int64); pmap[pose_ids] = np.arange(len(pose_ids))
    xmap = -np.ones(len(g.points), np.int64); xmap[point_ids] = np.arange(len(point_ids))
    omap = -np.ones(len(g.objects), np.int64); omap[obj_ids] = np.arange(len(obj_ids))

    s = synth.FactorGraph()
    s.cams = g.cams
    s.poses = np.ascontiguousarray(g.poses[pose_ids]); s.points = np.ascontiguousarray(g.points[point_ids])
    s.objects = np.ascontiguousarray(g.objects[obj_ids])
    s.reproj = dict(pose=pmap[rp["pose"][rp_idx]], point=xmap[rp["point"][rp_idx]], cam=rp["cam"][rp_idx], px=np.ascontiguousarray(rp["px"][rp_idx]),
                    sigma=rp["sigma"][rp_idx], huber=rp["huber"])
    s.bbox = dict(obj=omap[bb["obj"][bb_idx]], pose=pmap[bb["pose"][bb_idx]], cam=bb["cam"][bb_idx], corners=np.ascontiguousarray(bb["corners"][bb_idx]),
                  cov=np.ascontiguousarray(bb["cov"][bb_idx]), huber=bb["huber"], invalid_err=bb["invalid_err"])
    s.shape = dict(obj=omap[sh["obj"][sh_idx]], mean=sh["mean"][sh_idx], cov=sh["cov"][sh_idx], huber=sh["huber"])
    s.ltm = dict(obj=omap[lt["obj"][lt_idx]], mean=lt["mean"][lt_idx], cov=lt["cov"][lt_idx], huber=lt["huber"])
    s.relpose = dict(p1=pmap[rl["p1"][rl_idx]], p2=pmap[rl["p2"][rl_idx]], t=rl["t"][rl_idx], Rm=rl["Rm"][rl_idx], cov=rl["cov"][rl_idx], huber=rl["huber"])
    # constants (object_pose_graph_optimizer.h:440-472): frame 0 when the window starts there, else the leading frames
    nc = p.poses_prior_to_window_to_keep_constant if n_const is None else n_const
    s.const_pose = np.zeros(len(pose_ids), bool)
    if fix_poses:
        s.const_pose[:] = True
    elif start == 0:
        s.const_pose[0] = True
    else:
        s.const_pose[:max(1, nc)] = True
    s.const_point = np.zeros(len(point_ids), bool)
    s.const_obj = np.full(len(obj_ids), bool(fix_objects))
    return s, dict(pose=pose_ids, point=point_ids, obj=obj_ids, rp=rp_idx, bb=bb_idx)


def write_back(g, s, maps):
    g.poses[maps["pose"]] = s.poses; g.points[maps["point"]] = s.points; g.objects[maps["obj"]] = s.objects


def poses_stable(before, after, p: ScheduleParams):
    """isConsecutivePosesStable_ (offline_problem_runner.h:337-374): relative pose of consecutive frames may not jump."""
    if len(before) < 2:
        return True
    r0 = relative(before[:-1], before[1:]); r1 = relative(after[:-1], after[1:])
    d = relative(r0, r1)
    return bool(np.all(np.linalg.norm(d[:, :3], axis=1) <= p.consecutive_pose_transl_tol) and
                np.all(np.linalg.norm(d[:, 3:], axis=1) <= p.consecutive_pose_orient_tol))


# ----------------------------------------------------------------------------- backends
class GpuBackend:
    """The CUDA path: one Problem per window; outliers ranked on the device and excluded in place."""

    def __init__(self, ob, device=0):
        self.ob, self.device = ob, device
        self.stats = dict(solves=0, lm_steps=0, device_s=0.0, wall_s=0.0, structure_builds=0, excluded=0)

    def _acc(self, s):
        self.stats["solves"] += 1; self.stats["lm_steps"] += s.num_lm_steps; self.stats["device_s"] += s.minimizer_device_time_in_seconds
        return s

    def solve(self, sub, opts):
        t = time.time()
        p = self.ob.problem_from_graph(sub, device=self.device)
        s = self._acc(p.solve(**opts))
        self.stats["structure_builds"] += p.num_structure_builds(); self.stats["wall_s"] += time.time() - t
        return [s.final_cost]

    def two_phase(self, sub, opts1, opts2, frac):
        t = time.time()
        x0 = (sub.poses.copy(), sub.points.copy(), sub.objects.copy())
        p = self.ob.problem_from_graph(sub, device=self.device)
        s1 = self._acc(p.solve(**opts1))
        out = []
        if len(sub.reproj["pose"]):
            out += list(p.topk_outliers(self.ob.FACTOR_REPROJECTION, frac))
        if len(sub.bbox["obj"]):
            out += list(p.topk_outliers(self.ob.FACTOR_BBOX, frac))
        p.remove_residual_blocks(out)
        sub.poses[:], sub.points[:], sub.objects[:] = x0        # setValuesFromAnotherPoseGraph
        s2 = self._acc(p.solve(**opts2))
        self.stats["structure_builds"] += p.num_structure_builds(); self.stats["excluded"] += len(out)
        self.stats["wall_s"] += time.time() - t
        return [s1.final_cost, s2.final_cost]


# ----------------------------------------------------------------------------- the runner
def run_schedule(g, backend, p: ScheduleParams, max_frame=None, log=None):
    """OfflineProblemRunner::runOptimization (offline_problem_runner.h:100-270) on a synthetic session.

    `g` holds the session as the front end would deliver it: initial (odometry-integrated) poses, points / objects
    initialised relative to their first-observing keyframe, every consecutive rel-pose factor (the scope rule picks the
    ones it needs).  Frames are revealed one at a time; g is updated in place.  Returns the per-window cost log."""
    K = len(g.poses)
    max_frame = K - 1 if max_frame is None else max_frame
    init = g.poses.copy()
    rp, bb = g.reproj, g.bbox
    first_pt = np.full(len(g.points), K, np.int64); np.minimum.at(first_pt, rp["point"], rp["pose"])
    first_ob = np.full(len(g.objects), K, np.int64)
    if len(bb["obj"]):
        np.minimum.at(first_ob, bb["obj"], bb["pose"])
    pts_init, obj_init = g.points.copy(), g.objects.copy()
    out = []

    def add_frame(k):
        # pose chained from the previous optimised pose with the odometry increment (pose_graph_frame_data_adder.h:199-207)
        g.poses[k] = compose(g.poses[k - 1], relative(init[k - 1], init[k]))
        # map entities first seen from this frame are initialised relative to its current estimate
        for ids, first, x_init, arr in ((np.nonzero(first_pt == k)[0], first_pt, pts_init, g.points), (np.nonzero(first_ob == k)[0], first_ob, obj_init, g.objects)):
            if not len(ids):
                continue
            Ri, ti = _Rt(init[k]); Rc, tc = _Rt(g.poses[k])
            loc = (x_init[ids, 0:3] - ti) @ Ri            # R_i^T (x - t_i)
            arr[ids, 0:3] = loc @ Rc.T + tc
            if arr.shape[1] == 7:
                arr[ids, 3] = x_init[ids, 3] + synth.mat_to_rotvec(Rc @ Ri.T)[2]

    def visual_ba(start, nxt, ph1, ph2, tag):
        sub, maps = build_scope(g, start, nxt, p)
        before = sub.poses.copy()
        snap = (sub.poses.copy(), sub.points.copy(), sub.objects.copy())
        backend.kind = tag
        costs = backend.two_phase(sub, ph1.as_dict(), ph2.as_dict(), p.feature_outlier_percentage) if p.two_phase else backend.solve(sub, ph1.as_dict())
        if p.allow_reversion_after_detecting_jumps and not poses_stable(before, sub.poses, p):
            sub.poses[:], sub.points[:], sub.objects[:] = snap
            costs = costs + ["reverted"]
        write_back(g, sub, maps)
        out.append(dict(frame=nxt, start=start, kind=tag, costs=costs, n_reproj=len(maps["rp"]), n_bbox=len(maps["bb"])))
        if log:
            log(out[-1])

    def pgo_step(nxt, final):
        # tracking solve: only the newest pose moves (offline_problem_runner.h:438-496)
        nc = p.poses_prior_to_window_to_keep_constant
        sub, maps = build_scope(g, max(0, nxt - nc), nxt, p, n_const=nc)
        if not sub.const_pose.all():
            backend.kind = "tracking"
            c = backend.solve(sub, p.pre_pgo_tracking.as_dict()); write_back(g, sub, maps)
            out.append(dict(frame=nxt, start=int(maps["pose"][0]), kind="tracking", costs=c))
        # PGO with objects: rel-pose factor on every consecutive pair from the CURRENT estimates
        # (pose_graph_plus_objects_optimizer.h:94-159), bbox + shape + LTM factors, no visual factors
        rel_pts = None
        if p.enable_visual_non_opt_feature_adjustment_post_pgo:
            seen = np.nonzero(first_pt <= nxt)[0]
            R0, t0 = _Rt(g.poses[first_pt[seen]])
            rel_pts = (seen, np.einsum("nji,nj->ni", R0, g.points[seen] - t0))
        sub, maps = build_scope(g, 0, nxt, p, include_visual=False, relpose_mode="none", n_const=1)
        k = len(sub.poses)
        rel = relative(sub.poses[:-1], sub.poses[1:])
        Rm = synth.rotvec_to_mat(rel[:, 3:6])
        sub.relpose = dict(p1=np.arange(k - 1), p2=np.arange(1, k), t=np.ascontiguousarray(rel[:, :3]), Rm=Rm,
                           cov=synth.odom_cov(rel[:, :3], Rm, k=p.pgo_cov_multiplier), huber=p.pgo_huber)
        backend.kind = "pgo"
        c = backend.solve(sub, (p.final_pgo if final else p.pgo).as_dict()); write_back(g, sub, maps)
        out.append(dict(frame=nxt, start=0, kind="pgo", costs=c, n_bbox=len(maps["bb"])))
        if rel_pts is not None:   # re-anchor every point to its first-observing pose (:238-283)
            seen, loc = rel_pts
            R1, t1 = _Rt(g.poses[first_pt[seen]])
            g.points[seen] = np.einsum("nij,nj->ni", R1, loc) + t1
        if p.enable_visual_feats_only_opt_post_pgo:   # points-only BA (:285-353)
            sub, maps = build_scope(g, 0, nxt, p, include_objects=False, fix_poses=True, relpose_mode="none")
            if len(sub.points):
                backend.kind = "points_only"
                c = backend.solve(sub, p.post_pgo_vf_adjustment.as_dict()); write_back(g, sub, maps)
                out.append(dict(frame=nxt, start=0, kind="points_only", costs=c, n_reproj=len(maps["rp"])))
        if log:
            log(out[-1])

    def iteration(start, nxt, final):
        glob = is_global(start, nxt, p)
        run_pgo = glob and (p.use_pose_graph_on_final_global_ba if final else p.use_pose_graph_on_global_ba)
        run_vis = (not glob) or (p.use_visual_features_on_final_global_ba if final else p.use_visual_features_on_global_ba)
        if run_pgo:
            pgo_step(nxt, final)
        if run_vis:
            ph = (p.final_phase1, p.final_phase2) if final else ((p.gba_phase1, p.gba_phase2) if glob else (p.lba_phase1, p.lba_phase2))
            visual_ba(start, nxt, ph[0], ph[1], "final" if final else ("gba" if glob else "lba"))

    for nxt in range(1, max_frame + 1):
        add_frame(nxt)
        iteration(window_start(nxt, max_frame, p), nxt, False)
    iteration(0, max_frame, True)     # final refinement (offline_problem_runner.h:232-241)
    return out


class ShardedGpuBackend(GpuBackend):
    """GpuBackend for `world` ranks (one process per GPU): the large solves of the schedule -- kinds listed in `sharded_kinds`
    with at least `min_obs` reprojection + bounding-box blocks -- are sharded over the ranks (e-blocks dealt to ranks, one NCCL
    all-reduce of the reduced system per build, obvi_comm_*); everything else (local windows: too small to shard) is
    solved by rank 0 alone.  After every solve rank 0's parameter blocks are broadcast, so all ranks hold bit-identical state
    and take identical schedule decisions (windows are sequentially dependent: nothing runs concurrently).

    `dist` is torch.distributed (initialised, NCCL); `template` a Problem on this rank's device whose communicator
    (obvi_comm_init) the per-solve problems attach to."""

    def __init__(self, ob, device, dist, template, sharded_kinds=("pgo", "points_only", "gba", "final"), min_obs=200_000):
        super().__init__(ob, device)
        self.dist, self.template, self.sharded_kinds, self.min_obs = dist, template, set(sharded_kinds), min_obs
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.kind = None
        self.stats.update(sharded_solves=0, rank0_solves=0)

    def _bcast(self, sub):
        import torch
        for arr in (sub.poses, sub.points, sub.objects):
            if arr.size:
                t = torch.from_numpy(arr).cuda()
                self.dist.broadcast(t, 0)
                arr[...] = t.cpu().numpy()

    def _sharded(self, sub):
        return self.world > 1 and self.kind in self.sharded_kinds and len(sub.reproj["pose"]) + len(sub.bbox["obj"]) >= self.min_obs

    def _problem(self, sub, sharded):
        p = self.ob.problem_from_graph(sub, device=self.device)
        if sharded:
            p.comm_attach(self.template)
        return p

    def solve(self, sub, opts):
        t = time.time()
        sharded = self._sharded(sub)
        cost = [0.0]
        if sharded or self.rank == 0:
            p = self._problem(sub, sharded)
            s = self._acc(p.solve(**opts))
            cost = [s.final_cost]
            self.stats["structure_builds"] += p.num_structure_builds()
            self.stats["sharded_solves" if sharded else "rank0_solves"] += 1
        if self.world > 1:
            self._bcast(sub)
            cost = self._bcast_scalars(cost)
        self.stats["wall_s"] += time.time() - t
        return cost

    def _bcast_scalars(self, vals):
        import torch
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        self.dist.broadcast(t, 0)
        return t.tolist()

    def two_phase(self, sub, opts1, opts2, frac):
        t = time.time()
        sharded = self._sharded(sub)
        x0 = (sub.poses.copy(), sub.points.copy(), sub.objects.copy())
        costs = [0.0, 0.0]
        if not sharded:
            if self.rank == 0:
                costs = GpuBackend.two_phase(self, sub, opts1, opts2, frac)
                self.stats["rank0_solves"] += 2
            if self.world > 1:
                self._bcast(sub)
                costs = self._bcast_scalars(costs)
            return costs
        import torch
        p = self._problem(sub, True)
        s1 = self._acc(p.solve(**opts1))
        # outlier ranking needs the residuals of the WHOLE graph in order of addition: rank 0 ranks on an unsharded problem at the
        # phase-I solution and broadcasts the indices (the ranking itself is the device top-k of obvi_topk_outliers)
        n_rp, n_bb = len(sub.reproj["pose"]), len(sub.bbox["obj"])
        if self.rank == 0:
            q = self.ob.problem_from_graph(sub, device=self.device)
            out_rp = q.topk_outliers(self.ob.FACTOR_REPROJECTION, frac) if n_rp else np.zeros(0, np.uint64)
            out_bb = q.topk_outliers(self.ob.FACTOR_BBOX, frac) if n_bb else np.zeros(0, np.uint64)
            # factor ids are (type << 56 | index in order of addition): the same on every rank
            idx = np.concatenate([np.asarray(out_rp, np.uint64), np.asarray(out_bb, np.uint64)]).astype(np.int64)
            n = torch.tensor([len(idx)], dtype=torch.int64, device="cuda")
        else:
            n = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.dist.broadcast(n, 0)
        buf = torch.from_numpy(idx).cuda() if self.rank == 0 else torch.zeros(int(n.item()), dtype=torch.int64, device="cuda")
        if int(n.item()):
            self.dist.broadcast(buf, 0)
        p.remove_residual_blocks(buf.cpu().numpy().astype(np.uint64))
        sub.poses[:], sub.points[:], sub.objects[:] = x0
        s2 = self._acc(p.solve(**opts2))
        self.stats["structure_builds"] += p.num_structure_builds(); self.stats["excluded"] += int(n.item())
        self.stats["sharded_solves"] += 2
        self._bcast(sub)      # (already identical on every rank after the merged write-back; keeps the invariant explicit)
        self.stats["wall_s"] += time.time() - t
        return [s1.final_cost, s2.final_cost]
