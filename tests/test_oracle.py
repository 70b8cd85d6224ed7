"""CPU: the C++ oracle (Ceres-semantics restatement) against the golden vectors of the independent NumPy oracle,
and the two oracles against each other on fresh seeded graphs."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, graph_from_npz, rel_err


def test_kat_golden_vs_cpp_oracle(ob, oracle):
    d = np.load(os.path.join(GOLDEN, "kat_all_factors.npz"))
    g = graph_from_npz(ob, d)
    ev = oracle.evaluate(g, apply_loss=False)
    for k in ("r_reproj", "jp_reproj", "jl_reproj", "r_bbox", "jo_bbox", "jp_bbox", "r_shape", "j_shape", "r_ltm", "j_ltm", "r_rel", "j1_rel", "j2_rel"):
        assert d["kat__" + k].shape == ev[k].shape, k
        assert rel_err(ev[k], d["kat__" + k]) < 1e-10, k
    assert abs(ev["cost"] - float(d["kat__cost_raw"])) < 1e-9 * float(d["kat__cost_raw"])
    assert abs(oracle.evaluate(g, apply_loss=True)["cost"] - float(d["kat__cost_loss"])) < 1e-9 * float(d["kat__cost_loss"])


def test_lm_golden_vs_cpp_oracle(ob, oracle):
    d = np.load(os.path.join(GOLDEN, "lm_tiny.npz"))
    g = graph_from_npz(ob, d)
    res = oracle.solve(g, max_num_iterations=8, function_tolerance=1e-6, initial_radius=100.0, max_radius=1e4, use_nonmonotonic_steps=True)
    cost = np.array([it["cost"] for it in res["iterations"]])
    assert res["termination"] == str(d["lm__termination"])
    assert len(cost) == len(d["lm__cost"])
    assert np.all(np.abs(cost - d["lm__cost"]) <= 1e-8 * d["lm__cost"])
    assert [it["successful"] for it in res["iterations"]] == list(d["lm__successful"])
    assert abs(res["final_cost"] - float(d["lm__final_cost"])) <= 1e-8 * float(d["lm__final_cost"])
    assert np.abs(g.poses - d["lm__poses"]).max() < 1e-7 and np.abs(g.points - d["lm__points"]).max() < 1e-6
    assert np.abs(g.objects - d["lm__objects"]).max() < 1e-6


@pytest.mark.parametrize("nonmono", [False, True])
def test_two_oracles_agree_on_lm(ob, oracle, nonmono):
    from oracle import py_oracle as po
    g = ob.synth.make_graph(K=6, P=30, O=2, seed=21, objects_on=True, relpose="all", n_const_poses=1, min_point_obs=3, min_obj_obs=4,
                            ltm_frac=1.0, min_bbox_px=10.0, min_parallax_deg=0.0)
    g1, g2 = g.copy(), g.copy()
    o = dict(max_num_iterations=6, function_tolerance=1e-6, initial_radius=1e4, max_radius=1e16, use_nonmonotonic_steps=nonmono)
    a = po.solve_lm_dense(g1, **o)
    b = oracle.solve(g2, **o)
    assert a["termination"] == b["termination"] and a["lm_steps"] == b["lm_steps"]
    for x, y in zip(a["iterations"], b["iterations"]):
        assert x["successful"] == y["successful"] and abs(x["cost"] - y["cost"]) <= 1e-8 * abs(x["cost"])
    assert np.abs(g1.poses - g2.poses).max() < 1e-7


def test_fixed_cost_and_constant_blocks(ob, oracle):
    """Blocks whose parameters are all constant only contribute fixed cost; nothing constant may move."""
    g = ob.synth.make_graph(K=10, P=120, O=0, seed=22, objects_on=False, relpose="all", n_const_poses=4, min_point_obs=3, min_parallax_deg=0.0)
    g.const_point[:40] = True
    before = (g.poses.copy(), g.points.copy())
    res = oracle.solve(g, max_num_iterations=5, initial_radius=100.0, max_radius=1e4)
    assert res["fixed_cost"] > 0.0
    assert np.array_equal(g.poses[:4], before[0][:4]) and np.array_equal(g.points[:40], before[1][:40])
    assert res["final_cost"] < res["initial_cost"]


def test_huber_and_sqrt_information(oracle):
    from oracle import py_oracle as po
    rng = np.random.default_rng(0)
    for n in (3, 4, 6, 7):
        A = rng.normal(size=(n, n)); cov = A @ A.T + n * np.eye(n)
        sq = oracle.sqrt_information(cov)
        assert np.allclose(sq @ sq, np.linalg.inv(cov), rtol=1e-10, atol=1e-12) and np.allclose(sq, sq.T)
        assert np.allclose(sq, po.sqrt_inv_spd(cov), rtol=1e-10, atol=1e-12)
    assert po.huber_rho(0.25, 1.0) == (0.25, 1.0, 0.0)
    rho, rho1, rho2 = po.huber_rho(9.0, 1.0)
    assert rho == pytest.approx(5.0) and rho1 == pytest.approx(1.0 / 3.0) and rho2 < 0
