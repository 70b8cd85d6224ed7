"""Long-term-map extraction on top of the backend: the host logic of
src/refactoring/long_term_map/long_term_object_map_extraction.cpp (+ include/.../long_term_object_map_extraction.h), mirrored so
that the covariance path of the backend (`obvi_object_covariances`) can be driven -- and parity-checked -- the way the reference
drives ceres::Covariance, including its rank-repair retry loop.

    far-feature filter                 long_term_object_map_extraction.cpp:117-180  (features farther than
                                       `far_feature_threshold` from every camera that saw them lose all their factors)
    problem for the extraction         :52-80 (every frame, shape-dimension priors excluded, LTM objects forced in) and
                                       :103-105, 190-197 (max_num_iterations = 0: an evaluation, not a solve)
    covariance of every ellipsoid      :362-440 + long_term_object_map_extraction.h:455-520 (IndependentEllipsoids: the (o, o)
                                       7x7 block of every object, mean = the current estimate)
    rank repair                        findRankDeficiencies :503-760, addPriorToProblemParams :762-926, retry loop :928-1062
                                       (kMaxJacobianExtractionRetries = 5, kRankDeficiencyColsBuffer = 50, .h:20-21)

The device work (evaluation, Jacobian export, Schur elimination + reduced solves of the covariance blocks, the ParameterPrior
factors) is the backend's; this file is bookkeeping.  Two deliberate differences from the reference, both documented where they
occur: the rank of J is found with a dense column-pivoted QR (SuiteSparseQR is not available here) under SPQR's default
tolerance, which limits that step to a few thousand columns -- beyond that the deficiency is taken as 0 and the 50-column
buffer plus the retries do the work, which is what the buffer is there for (:617-624); and constant blocks are skipped when
columns are mapped back to parameters (the reference walks every listed block at full width, :655-757, although
Problem::Evaluate drops the columns of constant blocks).
"""
from dataclasses import dataclass, field

import numpy as np

K_MAX_JACOBIAN_EXTRACTION_RETRIES = 5   # long_term_object_map_extraction.h:20
K_RANK_DEFICIENCY_COLS_BUFFER = 50      # long_term_object_map_extraction.h:21
_WIDTH = {"point": 3, "pose": 6, "obj": 7}


@dataclass
class LtmExtractionParams:
    """LongTermMapExtractionTunableParams (long_term_map_extraction_tunable_params.h:10-16; config/base7a_2_fallback.json:116-121)."""
    far_feature_threshold: float = 75.0
    min_col_norm: float = 5e-4            # read by the reference but superseded by the minimum non-problem column norm (:765-766)
    dense_rank_max_cols: int = 4000       # this mirror's limit for the dense rank detection (see the module docstring)


@dataclass
class LtmResult:
    ok: bool
    objects: list = field(default_factory=list)        # object indices, ascending
    means: np.ndarray = None                           # (n, 7) ellipsoid estimates (EllipsoidResults, .h:497-507)
    covariances: np.ndarray = None                     # (n, 7, 7) marginal covariance blocks
    retries: int = 0
    added_priors: list = field(default_factory=list)   # (kind, index, param_idx, mean, std) in insertion order
    rank_deficiencies: list = field(default_factory=list)
    far_points: np.ndarray = None


# ----------------------------------------------------------------------------------------------- problem for the extraction
def far_feature_points(g, threshold):
    """Mask of the points whose distance to EVERY observing camera centre exceeds `threshold`
    (long_term_object_map_extraction.cpp:117-180: min over the feature's reprojection factors of |t_world<-cam - X|)."""
    from .synth import rotvec_to_mat
    rp = g.reproj
    far = np.zeros(len(g.points), dtype=bool)
    if len(rp["pose"]) == 0:
        return far
    R_wr = rotvec_to_mat(g.poses[:, 3:6])
    t_ext = np.stack([np.asarray(c["t"], dtype=np.float64) for c in g.cams])
    centre = g.poses[rp["pose"], 0:3] + np.einsum("nij,nj->ni", R_wr[rp["pose"]], t_ext[rp["cam"]])   # combinePoses(pose, extrinsics).transl_
    dist = np.linalg.norm(centre - g.points[rp["point"]], axis=1)
    dmin = np.full(len(g.points), np.inf)
    np.minimum.at(dmin, rp["point"], dist)
    far[np.isfinite(dmin) & (dmin > threshold)] = True
    return far


def ltm_graph(g, far):
    """The factor graph the extraction evaluates: every frame, reprojection factors of far features dropped, shape-dimension
    priors dropped (`factor_types_to_exclude = {kShapeDimPriorFactorTypeId}`, :68-69), everything else kept.  Parameter arrays
    are shared with `g` (the reference works on a copy of the pose graph and never writes it: max_num_iterations = 0)."""
    h = g.copy()
    h.poses, h.points, h.objects = g.poses, g.points, g.objects
    keep = ~far[g.reproj["point"]] if len(g.reproj["point"]) else np.zeros(0, dtype=bool)
    h.reproj = {k: (v[keep] if isinstance(v, np.ndarray) and len(v) == len(keep) else v) for k, v in g.reproj.items()}
    h.shape = {k: (v[:0] if isinstance(v, np.ndarray) else v) for k, v in g.shape.items()}
    return h


def ordered_blocks(g, added_priors=()):
    """Parameter blocks in the order findRankDeficiencies lists them (:556-589): features by id, frames by id, objects by
    id -- only blocks that some residual of the problem (or an added prior) touches; constant blocks are skipped here."""
    pts = set(int(i) for i in g.reproj["point"])
    frames = set(int(i) for i in g.reproj["pose"]) | set(int(i) for i in g.bbox["pose"]) | set(int(i) for i in g.relpose["p1"]) | \
        set(int(i) for i in g.relpose["p2"])
    objs = set(int(i) for i in g.bbox["obj"]) | set(int(i) for i in g.shape["obj"]) | set(int(i) for i in g.ltm["obj"])
    for kind, index, *_ in added_priors:
        {"point": pts, "pose": frames, "obj": objs}[kind].add(int(index))
    out = [("point", i) for i in sorted(pts) if not g.const_point[i]]
    out += [("pose", i) for i in sorted(frames) if not g.const_pose[i]]
    out += [("obj", i) for i in sorted(objs) if not g.const_obj[i]]
    return out


# ----------------------------------------------------------------------------------------------- rank repair
def rank_deficiency_dense(rows, cols, vals, shape, max_cols):
    """getRankDeficiency (:207-359): columns - rank of the Jacobian.  SuiteSparseQR(SPQR_DEFAULT_TOL) there; here a dense
    column-pivoted QR with the same tolerance, 20 (m + n) eps max_j |J_:j|_2.  Returns None above `max_cols` columns."""
    m, n = shape
    if n == 0:
        return 0
    if n > max_cols:
        return None
    from scipy.linalg import qr
    J = np.zeros((m, n))
    for i in range(m):
        sl = slice(rows[i], rows[i + 1])
        J[i, cols[sl]] = vals[sl]
    tol = 20.0 * (m + n) * np.finfo(np.float64).eps * np.sqrt((J * J).sum(axis=0).max())
    R = qr(J, mode="r", pivoting=True)[0]
    d = np.abs(np.diag(R[:min(m, n), :]))
    return int(n - np.count_nonzero(d > tol))


def column_sq_norms(cols, vals):
    """`norm_for_cols` (:593-601): sum of squares per column, only for columns that own at least one stored entry."""
    out = {}
    for c, v in zip(cols.tolist(), vals.tolist()):
        out[c] = out.get(c, 0.0) + v * v
    return out


def select_deficient_columns(norm_for_cols, rank_deficiency):
    """:617-654 -- the (rank_deficiency + buffer) columns of smallest squared norm get a prior; the next one up supplies
    `min_non_prob_col_norm_`.  Returns (sorted column list, that norm).  (Ties are broken by column index here; the reference
    partial-sorts an unordered_map, so its tie order is unspecified.)"""
    count = min(int(rank_deficiency) + K_RANK_DEFICIENCY_COLS_BUFFER + 1, len(norm_for_cols))
    if count == 0:
        return [], 0.0
    smallest = sorted(norm_for_cols.items(), key=lambda kv: (kv[1], kv[0]))[:count]
    return sorted(c for c, _ in smallest[:-1]), smallest[-1][1]


def columns_to_parameters(blocks, columns):
    """:655-757 -- walk the ordered blocks, turning column numbers into (kind, index, parameter index within the block)."""
    out, col0, it = [], 0, iter(sorted(columns))
    nxt = next(it, None)
    for kind, index in blocks:
        w = _WIDTH[kind]
        while nxt is not None and nxt < col0 + w:
            out.append((kind, index, nxt - col0))
            nxt = next(it, None)
        col0 += w
        if nxt is None:
            break
    return out


class GpuLtmBackend:
    """The extraction problem on the device: obvi_solve with 0 iterations (:103-105), obvi_object_covariances,
    obvi_evaluate_jacobian (Problem::Evaluate with a CRSMatrix, :251-252, 591-598) and obvi_factor_add_param_prior (:817, 869, 916)."""

    def __init__(self, ob, g, device=0):
        self.ob, self.g = ob, g
        self.p = ob.problem_from_graph(g, device=device)
        self.arr = dict(pose=g.poses, point=g.points, obj=g.objects)

    def evaluate(self):
        return self.p.solve(max_num_iterations=0).initial_cost

    def covariances(self, objs):
        try:
            blk = [self.g.objects[o] for o in objs]
            return True, self.p.object_covariances(blk, blk)
        except self.ob.ObviError:
            return False, None           # ceres::Covariance::Compute returned false (:432-438)

    def jacobian(self, blocks):
        rows, cols, vals, shape, _ = self.p.evaluate_jacobian(True, None, [self.arr[k][i] for k, i in blocks])
        return rows, cols, vals, shape

    def add_param_prior(self, kind, index, param_idx, mean, std):
        self.p.add_parameter_prior(self.arr[kind][index], int(param_idx), float(mean), float(std))   # no loss function (nullptr, :820)


def extract_long_term_map(g, make_backend, params: LtmExtractionParams = None, log=None):
    """extractLongTermObjectMap with the IndependentEllipsoids extractor (long_term_object_map_extraction.h:455-520), i.e.
    extractCovarianceWithRankDeficiencyHandling (:928-1062) + the per-object block read-out.  `make_backend(graph)` builds the
    problem (GpuLtmBackend, or the oracle-backed twin in tests/helpers.py)."""
    params = params or LtmExtractionParams()
    far = far_feature_points(g, params.far_feature_threshold)
    h = ltm_graph(g, far)
    B = make_backend(h)
    B.evaluate()
    objs = sorted((set(int(o) for o in h.bbox["obj"]) | set(int(o) for o in h.ltm["obj"])) - set(np.nonzero(h.const_obj)[0].tolist()))
    res = LtmResult(ok=False, objects=objs, far_points=far)
    if not objs:
        return res
    ok, cov = B.covariances(objs)
    while not ok and res.retries < K_MAX_JACOBIAN_EXTRACTION_RETRIES:
        res.retries += 1
        blocks = ordered_blocks(h, res.added_priors)
        rows, cols, vals, shape = B.jacobian(blocks)
        norms = column_sq_norms(cols, vals)
        rd = rank_deficiency_dense(rows, cols, vals, shape, params.dense_rank_max_cols)
        res.rank_deficiencies.append(rd)
        fix, min_norm = select_deficient_columns(norms, 0 if rd is None else rd)
        if log:
            log(f"retry {res.retries}: rank deficiency {rd}, {len(fix)} columns below {min_norm:.3e}")
        if not fix:
            break                                               # "No rank deficient columns identified" (:648-651)
        col0 = _column_offsets(blocks)
        for kind, index, pidx in columns_to_parameters(blocks, fix):
            gap = min_norm - norms[col0[(kind, index)] + pidx]
            if not gap > 0.0:
                continue                                        # the reference would add a prior with std = inf: no effect
            mean = float({"pose": h.poses, "point": h.points, "obj": h.objects}[kind][index][pidx])
            std = 1.0 / np.sqrt(gap)                            # brings the column's squared norm up to min_norm (:808, 860, 907)
            B.add_param_prior(kind, index, pidx, mean, std)
            res.added_priors.append((kind, int(index), int(pidx), mean, float(std)))
        B.evaluate()
        ok, cov = B.covariances(objs)
    res.ok = bool(ok)
    if ok:
        res.means = np.stack([h.objects[o].copy() for o in objs])
        res.covariances = cov
    return res


def _column_offsets(blocks):
    out, c = {}, 0
    for k, i in blocks:
        out[(k, i)] = c
        c += _WIDTH[k]
    return out
