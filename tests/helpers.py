"""Shared test helpers."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def graph_from_npz(ob, d):
    g = ob.synth.FactorGraph()
    g.poses, g.points, g.objects = (np.ascontiguousarray(d[k]).copy() for k in ("poses", "points", "objects"))
    g.const_pose, g.const_point, g.const_obj = (d[k].copy() for k in ("const_pose", "const_point", "const_obj"))
    g.cams = [dict(intr=tuple(d["cam_intr"][c]), R=d["cam_R"][c].copy(), t=d["cam_t"][c].copy()) for c in range(len(d["cam_intr"]))]
    for name in ("reproj", "bbox", "shape", "ltm", "relpose"):
        dd = {}
        for k in d.files:
            if k.startswith(name + "__"):
                v = d[k]
                dd[k[len(name) + 2:]] = float(v) if v.ndim == 0 else v.copy()
        setattr(g, name, dd)
    return g


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / (1.0 + np.abs(b).max())) if a.size else 0.0


class OracleBackend:
    """Schedule backend (obvi-slam_b200/schedule.py) that runs every solve on the CPU oracle: the checker for the GPU backend."""

    def __init__(self, oracle):
        self.oracle = oracle

    @staticmethod
    def _o(opts):
        return dict(max_num_iterations=opts["max_num_iterations"], function_tolerance=opts["function_tolerance"],
                    gradient_tolerance=opts["gradient_tolerance"], parameter_tolerance=opts["parameter_tolerance"],
                    initial_radius=opts["initial_trust_region_radius"], max_radius=opts["max_trust_region_radius"],
                    use_nonmonotonic_steps=bool(opts["use_nonmonotonic_steps"]))

    def solve(self, sub, opts):
        return [self.oracle.solve(sub, **self._o(opts))["final_cost"]]

    @staticmethod
    def _topk(sq, frac):
        by_err = {}
        for i, e in enumerate(sq):            # std::map<double, id, greater>: equal keys overwrite
            by_err[e] = i
        order = sorted(by_err, reverse=True)
        return [by_err[e] for e in order[:int(len(order) * frac)]]

    def two_phase(self, sub, opts1, opts2, frac):
        x0 = (sub.poses.copy(), sub.points.copy(), sub.objects.copy())
        c1 = self.oracle.solve(sub, **self._o(opts1))["final_cost"]
        ev = self.oracle.evaluate(sub, apply_loss=False)
        keep_rp = np.ones(len(sub.reproj["pose"]), bool); keep_bb = np.ones(len(sub.bbox["obj"]), bool)
        if len(keep_rp):
            r = ev["r_reproj"]; keep_rp[self._topk(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1], frac)] = False
        if len(keep_bb):
            r = ev["r_bbox"]; keep_bb[self._topk(((r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1]) + r[:, 2] * r[:, 2]) + r[:, 3] * r[:, 3], frac)] = False
        s2 = sub.copy()
        s2.poses[:], s2.points[:], s2.objects[:] = x0
        for k in ("pose", "point", "cam", "px", "sigma"):
            s2.reproj[k] = s2.reproj[k][keep_rp]
        for k in ("obj", "pose", "cam", "corners", "cov"):
            s2.bbox[k] = s2.bbox[k][keep_bb]
        c2 = self.oracle.solve(s2, **self._o(opts2))["final_cost"]
        sub.poses[:], sub.points[:], sub.objects[:] = s2.poses, s2.points, s2.objects
        return [c1, c2]




class OracleLtmBackend:
    """Twin of obvi-slam_b200/ltm_extraction.py:GpuLtmBackend on the dense NumPy oracle (small graphs): evaluation, marginal
    covariance blocks with ceres::Covariance's failure on a rank-deficient Jacobian, block-structured CRS export in a given
    block order, ParameterPrior insertion.  The checker for the GPU backend of the long-term-map extraction."""

    def __init__(self, po, ltm, g):
        self.po, self.ltm, self.g = po, ltm, g
        if not hasattr(g, "prior"):
            g.prior = dict(kind=[], index=[], idx=[], mean=[], std=[])

    def evaluate(self):
        return self.po.evaluate(self.g, want_jac=False, apply_loss=True)[0]

    def _dense(self):
        off, n = self.po._layout(self.g)
        _, _, J = self.po.evaluate(self.g, apply_loss=True)
        return off, n, J[:, :n]

    def _structure(self, J, off):
        """Stored-entry mask: a residual row stores every column of a block it touches (Ceres' CRS is block structured)."""
        width = dict(pose=6, point=3, obj=7)
        mask = np.zeros(J.shape, bool)
        for (kind, _), o in off.items():
            w = width[kind]
            mask[:, o:o + w] = (J[:, o:o + w] != 0).any(axis=1)[:, None]
        return mask

    def covariances(self, objs):
        off, n, J = self._dense()
        used = self._structure(J, off).any(axis=0)
        Ju = J[:, used]
        m, k = Ju.shape
        rows = np.arange(0, (m + 1) * k, k); cols = np.tile(np.arange(k), m)
        if self.ltm.rank_deficiency_dense(rows, cols, Ju.ravel(), (m, k), 10 ** 9) > 0:
            return False, None
        cov = np.zeros((n, n))
        cov[np.ix_(used, used)] = np.linalg.inv(Ju.T @ Ju)
        return True, np.stack([cov[off[("obj", o)]:off[("obj", o)] + 7, off[("obj", o)]:off[("obj", o)] + 7] for o in objs])

    def jacobian(self, blocks):
        off, n, J = self._dense()
        mask = self._structure(J, off)
        width = dict(pose=6, point=3, obj=7)
        sel = np.concatenate([np.arange(off[b], off[b] + width[b[0]]) for b in blocks]) if blocks else np.zeros(0, int)
        Jb, Mb = J[:, sel], mask[:, sel]
        rows = np.concatenate([[0], np.cumsum(Mb.sum(axis=1))]).astype(np.int64)
        r, c = np.nonzero(Mb)
        return rows, c.astype(np.int64), Jb[r, c], Jb.shape

    def add_param_prior(self, kind, index, param_idx, mean, std):
        pr = self.g.prior
        pr["kind"].append(kind); pr["index"].append(int(index)); pr["idx"].append(int(param_idx)); pr["mean"].append(float(mean)); pr["std"].append(float(std))
