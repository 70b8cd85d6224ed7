"""Generates tests/golden/*.npz from the INDEPENDENT NumPy oracle (oracle/py_oracle.py: closed-form residuals,
complex-step Jacobians, dense-normal-equation LM).  The reference holds no golden vectors for this path
(SURVEY.md section 8c), so these are the project's own pins; the C++ oracle and the CUDA path are checked
against them.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import obvi_b200 as ob  # noqa: E402
from oracle import py_oracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = {"reproj": ("r_reproj", ["jp_reproj", "jl_reproj"]), "bbox": ("r_bbox", ["jo_bbox", "jp_bbox"]), "shape": ("r_shape", ["j_shape"]),
         "ltm": ("r_ltm", ["j_ltm"]), "relpose": ("r_rel", ["j1_rel", "j2_rel"])}


def graph_arrays(g):
    d = dict(poses=g.poses, points=g.points, objects=g.objects, const_pose=g.const_pose, const_point=g.const_point, const_obj=g.const_obj,
             cam_intr=np.array([c["intr"] for c in g.cams]), cam_R=np.array([c["R"] for c in g.cams]), cam_t=np.array([c["t"] for c in g.cams]))
    for name in ("reproj", "bbox", "shape", "ltm", "relpose"):
        for k, v in getattr(g, name).items():
            d[f"{name}__{k}"] = np.asarray(v)
    return d


def kat(g):
    out = {}
    acc = {}
    for kind, a, fun, refs in po.residual_blocks(g):
        vals = [np.array(po._get(g, k, i), dtype=np.float64) for k, i in refs]
        r, J = po.complex_step_jacobian(fun, vals)
        rn, jn = KINDS[kind]
        acc.setdefault(rn, []).append(r)
        for name, Jk in zip(jn, J):
            acc.setdefault(name, []).append(Jk)
    for k, v in acc.items():
        out[k] = np.array(v)
    out["cost_raw"] = po.evaluate(g, want_jac=False, apply_loss=False)[0]
    out["cost_loss"] = po.evaluate(g, want_jac=False, apply_loss=True)[0]
    return out


def main():
    # 1. known-answer vectors: every factor type, incl. adversarial poses around the 1e-8 small-angle branch and near pi
    g = ob.synth.make_graph(K=8, P=60, O=3, seed=11, objects_on=True, relpose="all", n_const_poses=1, min_point_obs=3, min_obj_obs=4,
                            ltm_frac=0.5, min_bbox_px=10.0, min_parallax_deg=0.0)
    g.poses[2, 3:6] = [3e-9, 0.0, 0.0]
    g.poses[3, 3:6] = [2e-8, 1e-9, 0.0]
    g.poses[4, 3:6] = 0.0
    g.poses[5, 3:6] = np.array([0.6, 0.5, 0.62]) / np.linalg.norm([0.6, 0.5, 0.62]) * (np.pi - 1e-3)
    d = graph_arrays(g)
    d.update({f"kat__{k}": v for k, v in kat(g).items()})
    np.savez_compressed(os.path.join(HERE, "kat_all_factors.npz"), **d)
    print("kat:", g.counts())
    # 2. LM trajectory on a tiny well-posed graph (dense solve): cost per iteration + final parameters
    g = ob.synth.make_graph(K=6, P=40, O=2, seed=12, objects_on=True, relpose="all", n_const_poses=1, min_point_obs=3, min_obj_obs=4,
                            ltm_frac=1.0, min_bbox_px=10.0, min_parallax_deg=0.0)
    d = graph_arrays(g)
    g2 = g.copy()
    opts = dict(max_num_iterations=8, function_tolerance=1e-6, initial_radius=100.0, max_radius=1e4, use_nonmonotonic_steps=True)
    res = po.solve_lm_dense(g2, **opts)
    d["lm__cost"] = np.array([it["cost"] for it in res["iterations"]])
    d["lm__successful"] = np.array([it["successful"] for it in res["iterations"]])
    d["lm__step_norm"] = np.array([it["step_norm"] for it in res["iterations"]])
    d["lm__final_cost"] = res["final_cost"]
    d["lm__poses"], d["lm__points"], d["lm__objects"] = g2.poses, g2.points, g2.objects
    d["lm__termination"] = np.array(res["termination"])
    np.savez_compressed(os.path.join(HERE, "lm_tiny.npz"), **d)
    print("lm:", g.counts(), res["termination"], [f"{c:.6e}" for c in d["lm__cost"]])


if __name__ == "__main__":
    main()
