"""The C++ host side: header-only Ceres-compatibility shim + replacement factor headers (obvi-slam_b200/host/include).

CPU: the harness that restates the reference's call patterns (tests/cpp/shim_harness.cpp) compiles against the shim.
GPU: it is run on a dumped graph and must reproduce what the same sequence gives through the Python binding of the C ABI
(two-phase pattern: Solve, Evaluate raw residuals, drop the worst 10 % reprojection blocks, Solve again).
"""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def build_harness():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp")], stdout=subprocess.DEVNULL)
    return os.path.join(HERE, "cpp", "shim_harness")


def dump_graph(g, path, iters):
    n = g.counts()
    f = lambda a: np.asarray(a, dtype=np.float64).ravel()
    rp, bb, sh, lt, rl = g.reproj, g.bbox, g.shape, g.ltm, g.relpose
    head = [n["poses"], n["points"], n["objects"], len(g.cams), n["reproj"], n["bbox"], n["shape"], n["ltm"], n["relpose"],
            rp["huber"], bb["huber"], bb["invalid_err"], sh["huber"], lt["huber"], rl["huber"], iters]
    parts = [f(head), f(g.poses), f(g.points), f(g.objects), f(g.const_pose), f([c["intr"] for c in g.cams]), f([c["R"] for c in g.cams]),
             f([c["t"] for c in g.cams])]
    parts.append(f(np.column_stack([rp["pose"], rp["point"], rp["cam"], rp["px"], rp["sigma"]])) if n["reproj"] else f([]))
    parts.append(f(np.column_stack([bb["obj"], bb["pose"], bb["cam"], bb["corners"], bb["cov"].reshape(-1, 16)])) if n["bbox"] else f([]))
    parts.append(f(np.column_stack([sh["obj"], sh["mean"], sh["cov"].reshape(-1, 9)])) if n["shape"] else f([]))
    parts.append(f(np.column_stack([lt["obj"], lt["mean"], lt["cov"].reshape(-1, 49)])) if n["ltm"] else f([]))
    parts.append(f(np.column_stack([rl["p1"], rl["p2"], rl["t"], rl["Rm"].reshape(-1, 9), rl["cov"].reshape(-1, 36)])) if n["relpose"] else f([]))
    np.concatenate(parts).tofile(path)


def test_shim_harness_compiles():
    assert os.path.exists(build_harness())


@pytest.mark.gpu
def test_shim_matches_c_abi_two_phase(ob, tmp_path):
    exe = build_harness()
    g = ob.synth.make_graph(K=14, P=400, O=5, seed=31, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4, ltm_frac=0.5)
    iters = 10
    gin, gout = str(tmp_path / "graph.bin"), str(tmp_path / "result.bin")
    dump_graph(g, gin, iters)
    subprocess.check_call([exe, gin, gout])
    res = np.fromfile(gout)
    # the same sequence through the Python binding
    p = ob.problem_from_graph(g)
    o = dict(max_num_iterations=iters, function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8,
             initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)
    s1 = p.solve(**o)
    raw_cost, r = p.evaluate(apply_loss_function=False)
    out = p.topk_outliers(ob.FACTOR_REPROJECTION, 0.1)
    for fid in out:
        p.remove_residual_block(fid)
    s2 = p.solve(**o)
    head = res[:11]
    # The harness registers one heap block per parameter in order of first use, the binding registers whole arrays: the
    # internal point order (and with it the floating-point summation order) differs, so costs agree to rounding at
    # iteration 0 and to ~1e-6 after ten LM iterations of this deliberately noisy little problem.
    tol = lambda a, b, t: abs(a - b) <= t * max(1.0, abs(b))
    assert tol(head[0], s1.initial_cost, 1e-11) and tol(head[1], s1.final_cost, 1e-6) and int(head[2]) == s1.num_iterations
    assert tol(head[3], raw_cost, 1e-6) and int(head[4]) == len(r) and int(head[5]) == len(out) and len(out) > 0
    assert tol(head[6], s2.initial_cost, 1e-6) and tol(head[7], s2.final_cost, 1e-6) and int(head[8]) == s2.num_iterations
    assert int(head[9]) == p.num_residual_blocks() and int(head[10]) == s1.num_parameters_reduced
    n = g.counts()
    K, P = n["poses"], n["points"]
    # ceres::Covariance through the shim vs the binding (diagonal blocks of the ellipsoids registered with the problem)
    ncov = int(res[11])
    used = sorted(set(int(o) for o in g.bbox["obj"]) | set(int(o) for o in g.shape["obj"]) | set(int(o) for o in g.ltm["obj"]))
    assert ncov == len(used) > 0
    cov_shim = res[12:12 + 49 * ncov].reshape(ncov, 7, 7)
    cov_py = p.object_covariances([g.objects[o] for o in used], [g.objects[o] for o in used])
    for a, b in zip(cov_shim, cov_py):
        assert np.abs(a - b).max() <= 1e-3 * np.abs(np.diag(b)).max()
    # trailing five values: the callback section (terminate successfully at iteration 2, state updated every iteration)
    term3, n_it3, calls3, moved3, usable3 = res[-5:]
    assert int(term3) == 3 and int(n_it3) == 3 and int(calls3) == 3 and moved3 == 1.0 and usable3 == 1.0      # USER_SUCCESS
    res = np.concatenate([res[:11], res[12 + 49 * ncov:-5]])
    assert np.abs(res[11:11 + 6 * K].reshape(K, 6) - g.poses).max() < 1e-5
    assert np.abs(res[11 + 6 * K:11 + 6 * K + 3 * P].reshape(P, 3) - g.points).max() < 1e-3
    assert np.abs(res[11 + 6 * K + 3 * P:].reshape(-1, 7) - g.objects).max() < 1e-3
    assert s2.final_cost < s1.final_cost  # outliers removed
