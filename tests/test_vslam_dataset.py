"""The reference's own simulated sequences (data/vslam_set4, vslam_set7, vslam_superset1; format: data/vslam_set4/README.md) as
test inputs.  The reference holds no expected residuals / costs for the optimisation path, but these sequences are
(pose, landmark, keypoint) triples produced by its authors' simulator, i.e. known answers for the projection model of
ReprojectionCostFunctor (reprojection_cost_functor.h:56-93 -> getProjectedPixelLocation, vslam_math_util.h:347-411): on the
noise-free sets the reprojection residual must vanish.  Fixtures: tests/golden/vslam_*.npz (tests/golden/make_vslam_fixtures.py)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, graph_from_npz

REF_DATA = "/root/reference/data"


def _fixture(ob, name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return graph_from_npz(ob, d), d


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("name,rel", [("vslam_set7", "vslam_set7"), ("vslam_superset1_low_density_high_noise", "vslam_superset1/low_density/high_noise")])
def test_reader_reproduces_the_committed_fixture(ob, name, rel):
    g, frame_ids, feature_ids = ob.vslam_dataset_io.read_vslam_dataset(os.path.join(REF_DATA, rel))
    f, d = _fixture(ob, name)
    assert list(d["frame_ids"]) == frame_ids and list(d["feature_ids"]) == feature_ids
    for a, b in ((g.poses, f.poses), (g.points, f.points), (g.reproj["px"], f.reproj["px"]), (g.reproj["pose"], f.reproj["pose"]),
                 (g.reproj["point"], f.reproj["point"])):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["vslam_set7", "vslam_set4"])
def test_noise_free_sequences_reproject_onto_their_keypoints(ob, oracle, name):
    """Both oracles, raw residuals (no loss): |r| * sigma is the pixel error.  The files carry 6 decimals; 5e-4 px is their noise
    floor.  vslam_set7 contains a frame rotated by exactly pi and one by -pi/2 (angle-axis edge of the pose parameterisation)."""
    from oracle import py_oracle as po
    g, _ = _fixture(ob, name)
    assert np.abs(np.linalg.norm(g.poses[:, 3:6], axis=1) - np.pi).min() < 1e-6 or name != "vslam_set7"
    sigma = g.reproj["sigma"][0]
    r_cpp = oracle.evaluate(g, apply_loss=False)["r_reproj"] * sigma
    assert r_cpp.shape == (len(g.reproj["pose"]), 2) and np.abs(r_cpp).max() < 5e-4
    # the NumPy oracle on a sample of the blocks (it evaluates block by block)
    idx = np.arange(0, len(g.reproj["pose"]), max(1, len(g.reproj["pose"]) // 300))
    for n in idx:
        c = g.cams[g.reproj["cam"][n]]
        r = po.reproj_residual(g.poses[g.reproj["pose"][n]], g.points[g.reproj["point"][n]], g.reproj["px"][n], c["intr"], c["R"], c["t"], sigma)
        assert np.abs(r * sigma).max() < 5e-4 and np.abs(r * sigma - r_cpp[n]).max() < 1e-9


def test_bundle_adjustment_of_a_reference_sequence_numpy_vs_cpp_oracle(ob, oracle):
    """vslam_superset1 (41 frames on a sine path, noisy keypoints): the two first poses constant (gauge + scale of the monocular
    problem), the others and the landmarks perturbed; the LM trajectories of the dense NumPy oracle and of the C++ restatement
    (Schur + sparse Cholesky) must coincide."""
    from oracle import py_oracle as po
    g, _ = _fixture(ob, "vslam_superset1_low_density_high_noise")
    rng = np.random.default_rng(5)
    g.const_pose[:2] = True
    g.poses[2:, 0:3] += rng.normal(0, 0.05, (len(g.poses) - 2, 3))
    g.poses[2:, 3:6] += rng.normal(0, 0.01, (len(g.poses) - 2, 3))
    g.points += rng.normal(0, 0.1, g.points.shape)
    g2 = g.copy()
    opts = dict(max_num_iterations=4, function_tolerance=1e-6, initial_radius=100.0, max_radius=1e4, use_nonmonotonic_steps=True)
    a = po.solve_lm_dense(g, **opts)
    b = oracle.solve(g2, **opts)
    assert len(a["iterations"]) == len(b["iterations"]) and a["iterations"][-1]["cost"] < 0.6 * a["iterations"][0]["cost"]
    for x, y in zip(a["iterations"], b["iterations"]):
        assert x["successful"] == y["successful"] and abs(x["cost"] - y["cost"]) <= 1e-9 * abs(y["cost"])
    assert np.abs(g.poses - g2.poses).max() < 1e-7 and np.abs(g.points - g2.points).max() < 1e-6


@pytest.mark.gpu
def test_gpu_bundle_adjustment_of_a_reference_sequence(ob, oracle):
    """The dense variant of the same sequence (41 frames, 937 landmarks, 9798 keypoints, one camera, tracks of up to 41 frames --
    longer than the row-owner kernels' span, so those landmarks take the generic elimination path) through the C ABI against
    the C++ oracle: same accept / reject sequence, cost per iteration within 1e-5, translations within 1e-4."""
    g, _ = _fixture(ob, "vslam_superset1_high_density_high_noise")
    rng = np.random.default_rng(6)
    g.const_pose[:2] = True
    g.poses[2:, 0:3] += rng.normal(0, 0.05, (len(g.poses) - 2, 3))
    g.poses[2:, 3:6] += rng.normal(0, 0.01, (len(g.poses) - 2, 3))
    g.points += rng.normal(0, 0.1, g.points.shape)
    g_ref = g.copy()
    p = ob.problem_from_graph(g)
    s = p.solve(max_num_iterations=8, function_tolerance=1e-6, initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)
    ref = oracle.solve(g_ref, max_num_iterations=8, function_tolerance=1e-6, initial_radius=100.0, max_radius=1e4, use_nonmonotonic_steps=True)
    msg = "\n".join(f"{a['iteration']:3d} gpu {a['cost']:.10e} ok={int(a['successful'])} | cpu {b['cost']:.10e} ok={int(b['successful'])}"
                    for a, b in zip(s.iterations, ref["iterations"]))
    print(msg)
    assert s.kernel_launches > 0 and s.num_iterations == len(ref["iterations"]), msg
    for a, b in zip(s.iterations, ref["iterations"]):
        assert a["successful"] == b["successful"], msg
        assert abs(a["cost"] - b["cost"]) <= 1e-5 * abs(b["cost"]), msg
    assert np.abs(g.poses[:, :3] - g_ref.poses[:, :3]).max() < 1e-4
