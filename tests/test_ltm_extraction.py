"""Long-term-map extraction mirror (obvi-slam_b200/ltm_extraction.py; reference:
src/refactoring/long_term_map/long_term_object_map_extraction.cpp): far-feature filter, the problem it evaluates, the
rank-repair bookkeeping -- on the CPU against the dense NumPy oracle, on the GPU against that same oracle run."""
import numpy as np
import pytest

from helpers import OracleLtmBackend


def _gauge_free_graph(ob, seed=32):
    # no constant pose, no rel-pose factors: the 6-dof gauge freedom makes J rank deficient (ceres::Covariance fails)
    return ob.synth.make_graph(K=8, P=150, O=3, seed=seed, objects_on=True, relpose="none", n_const_poses=0, min_obj_obs=4)


def test_far_feature_filter_and_extraction_graph(ob):
    ltm = ob.ltm_extraction
    g = ob.synth.make_graph(K=12, P=300, O=4, seed=7, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4)
    thr = 12.0
    far = ltm.far_feature_points(g, thr)
    # brute force, factor by factor (long_term_object_map_extraction.cpp:125-168)
    R = ob.synth.rotvec_to_mat(g.poses[:, 3:6])
    dmin = np.full(len(g.points), np.inf)
    for k, j, c in zip(g.reproj["pose"], g.reproj["point"], g.reproj["cam"]):
        centre = g.poses[k, :3] + R[k] @ np.asarray(g.cams[c]["t"])
        dmin[j] = min(dmin[j], np.linalg.norm(centre - g.points[j]))
    want = np.isfinite(dmin) & (dmin > thr)
    assert np.array_equal(far, want) and 0 < far.sum() < len(far)
    h = ltm.ltm_graph(g, far)
    assert h.counts()["shape"] == 0 and g.counts()["shape"] > 0                      # shape-dimension priors are excluded
    assert not far[h.reproj["point"]].any() and h.counts()["reproj"] == int((~far[g.reproj["point"]]).sum())
    assert h.counts()["bbox"] == g.counts()["bbox"] and h.counts()["relpose"] == g.counts()["relpose"]
    assert h.points is g.points                                                       # evaluation only: blocks are shared
    assert ltm.far_feature_points(g, 1e9).sum() == 0


def test_rank_repair_bookkeeping(ob):
    ltm = ob.ltm_extraction
    norms = {c: float(v) for c, v in enumerate([5.0, 1e-9, 3.0, 2e-9, 4.0, 0.5, 6.0])}
    saved = ltm.K_RANK_DEFICIENCY_COLS_BUFFER
    try:
        ltm.K_RANK_DEFICIENCY_COLS_BUFFER = 1
        fix, mn = ltm.select_deficient_columns(norms, 1)     # 1 + 1 columns get priors, the third-smallest is the reference norm
        assert fix == [1, 3] and mn == 0.5
        fix, mn = ltm.select_deficient_columns(norms, 50)    # more than there are: every column but the largest
        assert fix == [0, 1, 2, 3, 4, 5] and mn == 6.0
    finally:
        ltm.K_RANK_DEFICIENCY_COLS_BUFFER = saved
    blocks = [("point", 4), ("point", 9), ("pose", 2), ("obj", 1)]
    assert ltm.columns_to_parameters(blocks, [0, 5, 6, 11, 12, 18]) == [("point", 4, 0), ("point", 9, 2), ("pose", 2, 0), ("pose", 2, 5),
                                                                         ("obj", 1, 0), ("obj", 1, 6)]
    # dense rank detection: a 6 x 4 matrix with one dependent column and one zero column
    rng = np.random.default_rng(0)
    A = rng.normal(size=(6, 4)); A[:, 2] = A[:, 0] - 2 * A[:, 1]; A[:, 3] = 0
    rows = np.arange(0, 28, 4); cols = np.tile(np.arange(4), 6)
    assert ltm.rank_deficiency_dense(rows, cols, A.ravel(), (6, 4), 100) == 2
    assert ltm.rank_deficiency_dense(rows, cols, A.ravel(), (6, 4), 3) is None


def test_extraction_with_rank_repair_on_oracle(ob):
    """Gauge-free problem: the first covariance attempt fails, the repair adds (rank deficiency + 50) single-coordinate priors on
    the weakest columns, the retry succeeds (extractCovarianceWithRankDeficiencyHandling, :928-1062)."""
    from oracle import py_oracle as po
    ltm = ob.ltm_extraction
    g = _gauge_free_graph(ob)
    res = ltm.extract_long_term_map(g, lambda h: OracleLtmBackend(po, ltm, h))
    assert res.ok and res.retries == 1 and res.rank_deficiencies == [6]
    assert len(res.added_priors) == 6 + ltm.K_RANK_DEFICIENCY_COLS_BUFFER
    assert res.objects == sorted(set(int(o) for o in g.bbox["obj"]))
    assert np.array_equal(res.means, g.objects[res.objects])
    for c in res.covariances:
        assert np.allclose(c, c.T, rtol=1e-8, atol=1e-12) and np.all(np.linalg.eigvalsh(0.5 * (c + c.T)) > 0)
    # a well-posed problem needs no retry and no priors
    g2 = ob.synth.make_graph(K=8, P=150, O=3, seed=32, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4)
    res2 = ltm.extract_long_term_map(g2, lambda h: OracleLtmBackend(po, ltm, h))
    assert res2.ok and res2.retries == 0 and not res2.added_priors


@pytest.mark.gpu
def test_extraction_with_rank_repair_gpu_matches_oracle(ob):
    """The same extraction through the C ABI (0-iteration solve, obvi_object_covariances failing on the rank-deficient problem,
    obvi_evaluate_jacobian for the column norms / rank, obvi_factor_add_param_prior, retry).  Checked two ways: the repair
    bookkeeping against the oracle-backed run of the same extraction, and the covariance blocks against the dense inverse of
    J^T J of the graph WITH the priors the GPU run added (valid however many retries it took)."""
    from oracle import py_oracle as po
    ltm = ob.ltm_extraction
    g = _gauge_free_graph(ob)
    ref = ltm.extract_long_term_map(g.copy(), lambda h: OracleLtmBackend(po, ltm, h))
    res = ltm.extract_long_term_map(g, lambda h: ltm.GpuLtmBackend(ob, h), log=print)
    print("gpu: ok", res.ok, "retries", res.retries, "rank deficiencies", res.rank_deficiencies, "priors", len(res.added_priors))
    assert res.ok and 1 <= res.retries <= ltm.K_MAX_JACOBIAN_EXTRACTION_RETRIES
    assert res.rank_deficiencies[0] == ref.rank_deficiencies[0] == 6
    n1 = 6 + ltm.K_RANK_DEFICIENCY_COLS_BUFFER                      # first repair round: same columns, same weights as the oracle run
    assert [a[:3] for a in res.added_priors[:n1]] == [a[:3] for a in ref.added_priors[:n1]]
    assert np.allclose([a[4] for a in res.added_priors[:n1]], [a[4] for a in ref.added_priors[:n1]], rtol=1e-6)
    h = ltm.ltm_graph(g, res.far_points)
    h.prior = dict(kind=[a[0] for a in res.added_priors], index=[a[1] for a in res.added_priors], idx=[a[2] for a in res.added_priors],
                   mean=[a[3] for a in res.added_priors], std=[a[4] for a in res.added_priors])
    want = po.covariance_blocks(h, [(o, o) for o in res.objects])
    for a, b in zip(res.covariances, want):
        scale = np.sqrt(np.outer(np.abs(np.diag(b)), np.abs(np.diag(b))))
        print("covariance block: max |gpu - dense| / scale", np.abs(a - b).max() / scale.max())
        assert np.abs(a - b).max() <= 1e-4 * scale.max()
