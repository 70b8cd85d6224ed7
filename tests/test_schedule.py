"""BASELINE config 4: the reference's sliding-window, two-phase schedule (obvi-slam_b200/schedule.py mirrors
offline_problem_runner.h:100-958 + run_opt_utils.h:101-116 + pose_graph_plus_objects_optimizer.h:24-353).

CPU: window rule, scope selection and the composition on the CPU oracle.  GPU: the same session through the CUDA backend
(device-side outlier ranking + in-place exclusion) must follow the oracle's trajectory window by window."""
import numpy as np
import pytest

from helpers import OracleBackend


def small_params(ob, **kw):
    S = ob.schedule
    sp = lambda it, ftol: S.SolverParams(max_num_iterations=it, function_tolerance=ftol)
    p = S.ScheduleParams(local_ba_window_size=8, global_ba_frequency=6, poses_prior_to_window_to_keep_constant=2,
                         min_low_level_feature_observations=3, min_low_level_feature_observations_per_frame=20,
                         min_object_observations=4, feature_outlier_percentage=0.1,
                         lba_phase1=sp(6, 1e-3), lba_phase2=sp(8, 1e-4), gba_phase1=sp(8, 1e-6), gba_phase2=sp(8, 1e-6),
                         final_phase1=sp(8, 1e-6), final_phase2=sp(8, 1e-6), pgo=sp(8, 1e-6), final_pgo=sp(8, 1e-6),
                         pre_pgo_tracking=sp(5, 1e-3), post_pgo_vf_adjustment=sp(5, 1e-3))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def session(ob, K=20, seed=41):
    return ob.synth.make_graph(K=K, P=40 * K, O=5, seed=seed, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4)


def test_window_rule_matches_reference(ob):
    S = ob.schedule
    p = S.ScheduleParams()          # window 50, global every 30 (config/base7a_2_fallback.json:415-418)
    mx = 200
    assert [S.window_start(n, mx, p) for n in (1, 29, 30, 49, 50, 51, 59, 60, 61, 90, 199, 200)] == [0, 0, 0, 0, 0, 1, 9, 0, 11, 0, 149, 0]
    # frame 30 is a LOCAL BA over [0, 30]; 60, 90, ... and the last frame are global
    assert not S.is_global(0, 30, p) and S.is_global(0, 60, p) and S.is_global(0, 200, p) and not S.is_global(11, 61, p)


def test_scope_selection(ob):
    S = ob.schedule
    g = session(ob)
    p = small_params(ob)
    sub, maps = S.build_scope(g, 5, 13, p)
    assert len(sub.poses) == 9 and sub.const_pose[:2].all() and not sub.const_pose[2:].any()
    # every kept feature has >= 3 observations inside the window, every kept observation lies in it
    assert np.all((g.reproj["pose"][maps["rp"]] >= 5) & (g.reproj["pose"][maps["rp"]] <= 13))
    assert np.bincount(sub.reproj["point"]).min() >= 3
    # rel-pose factors only touch feature-starved frames
    per_frame = np.bincount(sub.reproj["pose"], minlength=9)
    assert all(per_frame[a] < 20 or per_frame[b] < 20 for a, b in zip(sub.relpose["p1"], sub.relpose["p2"]))
    sub0, _ = S.build_scope(g, 0, 7, p)
    assert sub0.const_pose[0] and not sub0.const_pose[1:].any()
    # PGO scope: no visual factors, all objects kept by the bbox-count rule
    subp, _ = S.build_scope(g, 0, 19, p, include_visual=False, relpose_mode="none", n_const=1)
    assert len(subp.reproj["pose"]) == 0 and len(subp.points) == 0 and len(subp.bbox["obj"]) > 0


def test_pose_algebra_roundtrip(ob):
    S = ob.schedule
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.normal(size=(50, 3)), rng.normal(scale=0.7, size=(50, 3))], axis=1)
    b = np.concatenate([rng.normal(size=(50, 3)), rng.normal(scale=0.7, size=(50, 3))], axis=1)
    assert np.abs(S.compose(a, S.relative(a, b)) - b).max() < 1e-12


def test_schedule_on_oracle_runs_every_step_kind(ob, oracle):
    S = ob.schedule
    g = session(ob, K=16, seed=42)
    log = S.run_schedule(g, OracleBackend(oracle), small_params(ob))
    kinds = [e["kind"] for e in log]
    assert kinds.count("lba") >= 8 and "pgo" in kinds and "points_only" in kinds and kinds[-1] == "final"
    assert all(len(e["costs"]) >= 2 and e["costs"][1] <= e["costs"][0] * 1.5 for e in log if e["kind"] in ("lba", "final"))
    # 16 keyframes drift by millimetres only: the optimised trajectory must stay at the noise floor of 1 px features
    assert np.linalg.norm(g.poses[:, :3] - g.poses_gt[:, :3], axis=1).max() < 0.05
    assert np.isfinite(g.points).all() and np.isfinite(g.objects).all()


@pytest.mark.gpu
def test_schedule_gpu_matches_oracle(ob, oracle):
    S = ob.schedule
    g_gpu = session(ob, K=20, seed=43)
    g_cpu = g_gpu.copy()
    p = small_params(ob)
    be = S.GpuBackend(ob)
    log_gpu = S.run_schedule(g_gpu, be, p)
    log_cpu = S.run_schedule(g_cpu, OracleBackend(oracle), p)
    assert [(e["frame"], e["start"], e["kind"]) for e in log_gpu] == [(e["frame"], e["start"], e["kind"]) for e in log_cpu]
    for a, b in zip(log_gpu, log_cpu):
        assert len(a["costs"]) == len(b["costs"])
        for x, y in zip(a["costs"], b["costs"]):
            assert x == y or abs(x - y) <= 1e-5 * abs(y), (a, b)
    assert np.abs(g_gpu.poses[:, :3] - g_cpu.poses[:, :3]).max() < 1e-4
    assert np.abs(g_gpu.objects - g_cpu.objects).max() < 1e-3
    # one structure build per window problem: the phase-II exclusion never rebuilt
    assert be.stats["excluded"] > 0 and be.stats["structure_builds"] == sum(1 for e in log_gpu)


@pytest.mark.gpu
def test_multi_session_ltm_chain_matches_oracle(ob, oracle):
    """BASELINE config 5 shape at test size: sessions chained through the long-term map
    (src/evaluation/ltm_trajectory_sequence_executor.py:44-83): solve session k, extract the ellipsoid estimates and their
    marginal covariances (IndependentEllipsoids extractor, long_term_object_map_extraction.h:455-520), hand them to
    session k + 1 as LTM prior factors (independent_object_map_factor.h:21-33).  GPU chain vs the same chain on the CPU
    oracles (C++ LM + NumPy dense covariance)."""
    from oracle import py_oracle as po
    opts = dict(max_num_iterations=12, function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8,
                initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=0)
    o_opts = OracleBackend._o(opts)

    def sessions():
        return [ob.synth.make_graph(K=14, P=350, O=5, seed=71, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4) for _ in range(3)]

    def chain(gs, solve, covariances):
        prior = None
        finals = []
        for g in gs:
            if prior is not None:
                objs, mean, cov = prior
                g.ltm = dict(obj=objs.copy(), mean=mean.copy(), cov=cov.copy(), huber=1.0)
                g.objects[objs] = mean          # the next session starts from the map
            finals.append(solve(g))
            objs = np.array(sorted(set(int(o) for o in g.bbox["obj"])), np.int64)
            prior = (objs, g.objects[objs].copy(), covariances(g, objs))
        return finals, prior

    def gpu_solve(g):
        p = ob.problem_from_graph(g)
        gpu_solve.p = p
        return p.solve(**opts).final_cost

    def gpu_cov(g, objs):
        return gpu_solve.p.object_covariances([g.objects[o] for o in objs], [g.objects[o] for o in objs])

    fa, pa = chain(sessions(), gpu_solve, gpu_cov)
    fb, pb = chain(sessions(), lambda g: oracle.solve(g, **o_opts)["final_cost"], lambda g, objs: po.covariance_blocks(g, [(int(o), int(o)) for o in objs]))
    for x, y in zip(fa, fb):
        assert abs(x - y) <= 1e-5 * abs(y)
    assert np.array_equal(pa[0], pb[0])
    assert np.abs(pa[1] - pb[1]).max() < 1e-4
    assert np.abs(pa[2] - pb[2]).max() <= 1e-4 * np.abs(pb[2]).max()
    # the map tightens from session to session
    assert np.trace(pa[2][0]) > 0
