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
