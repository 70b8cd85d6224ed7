"""obvi-slam_b200 -- B200-native bundle-adjustment backend for ObVi-SLAM (Python host mirror).

This module is a thin ctypes binding of the C ABI in include/obvi_ba.h (libobvi_ba.so, built from
csrc/ for sm_100a).  It mirrors the slice of the ceres::Problem / ceres::Solve API that the reference
uses: add parameter blocks, add residual blocks through the reference's factor
factories, set blocks constant, Solve, Evaluate.  There is NO CPU fallback: importing works anywhere,
but creating a Problem without the CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OBVI_LIB_PATH") or os.path.join(_HERE, "libobvi_ba.so")   # OBVI_LIB_PATH: A/B builds of the same library
_LIB = None

_d = C.POINTER(C.c_double)
_pp = C.POINTER(C.c_void_p)

FACTOR_REPROJECTION, FACTOR_BBOX, FACTOR_SHAPE_PRIOR, FACTOR_LTM_PRIOR, FACTOR_REL_POSE, FACTOR_PARAM_PRIOR = 0, 2, 3, 4, 5, 6
TERMINATION = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE", 3: "USER_SUCCESS", 4: "USER_FAILURE"}
ITERATION_CALLBACK = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p)


class ObviError(RuntimeError):
    pass


class SolverOptions(C.Structure):
    """obvi_solver_options (Solver::Options subset of object_pose_graph_optimizer.h:651-672)."""
    _fields_ = [("max_num_iterations", C.c_int32), ("use_nonmonotonic_steps", C.c_int32),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_consecutive_nonmonotonic_steps", C.c_int32), ("max_num_consecutive_invalid_steps", C.c_int32),
                ("pcg_max_iterations", C.c_int32), ("pcg_relative_tolerance", C.c_double),
                ("iteration_callback", ITERATION_CALLBACK), ("iteration_callback_user", C.c_void_p),
                ("update_state_every_iteration", C.c_int32)]

    def __init__(self, **kw):
        super().__init__()
        lib().obvi_solver_options_init(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class IterationSummary(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
                ("linear_solver_iterations", C.c_int32), ("cost", C.c_double), ("cost_change", C.c_double),
                ("gradient_max_norm", C.c_double), ("step_norm", C.c_double), ("relative_decrease", C.c_double),
                ("trust_region_radius", C.c_double)]


class Summary(C.Structure):
    _fields_ = [("termination_type", C.c_int32), ("num_iterations", C.c_int32), ("num_lm_steps", C.c_int32),
                ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
                ("num_parameter_blocks_reduced", C.c_int32), ("num_parameters_reduced", C.c_int32),
                ("num_residual_blocks_reduced", C.c_int32), ("num_residuals_reduced", C.c_int32),
                ("is_solution_usable", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("fixed_cost", C.c_double), ("total_time_in_seconds", C.c_double),
                ("preprocessor_time_in_seconds", C.c_double), ("linear_solver_time_in_seconds", C.c_double),
                ("jacobian_evaluation_time_in_seconds", C.c_double), ("residual_evaluation_time_in_seconds", C.c_double),
                ("minimizer_device_time_in_seconds", C.c_double), ("pcg_iterations_total", C.c_int64),
                ("kernel_launches", C.c_int64)]

    iterations: list = []

    @property
    def termination(self):
        return TERMINATION[self.termination_type]

    def IsSolutionUsable(self):
        return bool(self.is_solution_usable)

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["termination"] = self.termination
        return d


def lib():
    """Load libobvi_ba.so; raises ObviError when it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ObviError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). This backend has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    u64p = C.POINTER(C.c_uint64)
    i32p = C.POINTER(C.c_int32)
    sig = {
        "obvi_problem_create": ([C.c_int, C.POINTER(vp)], C.c_int),
        "obvi_problem_destroy": ([vp], None),
        "obvi_last_error": ([vp], C.c_char_p),
        "obvi_version": ([], C.c_char_p),
        "obvi_param_add": ([vp, vp, C.c_int], C.c_int),
        "obvi_param_add_array": ([vp, vp, C.c_int, i64], C.c_int),
        "obvi_param_remove": ([vp, vp], C.c_int),
        "obvi_param_set_constant": ([vp, vp, C.c_int], C.c_int),
        "obvi_param_is_constant": ([vp, vp, C.POINTER(C.c_int)], C.c_int),
        "obvi_camera_add": ([vp, _d, _d, _d, C.POINTER(C.c_int)], C.c_int),
        "obvi_factor_add_reproj": ([vp, vp, vp, C.c_int, _d, dbl, dbl, u64p], C.c_int),
        "obvi_factor_add_reproj_batch": ([vp, i64, vp, vp, i32p, _d, _d, dbl, u64p], C.c_int),
        "obvi_factor_add_bbox": ([vp, vp, vp, C.c_int, _d, _d, dbl, dbl, u64p], C.c_int),
        "obvi_factor_add_bbox_batch": ([vp, i64, vp, vp, i32p, _d, _d, dbl, dbl, u64p], C.c_int),
        "obvi_factor_add_shape_prior": ([vp, vp, _d, _d, dbl, u64p], C.c_int),
        "obvi_factor_add_ltm_prior": ([vp, vp, _d, _d, dbl, u64p], C.c_int),
        "obvi_factor_add_rel_pose": ([vp, vp, vp, _d, _d, _d, dbl, u64p], C.c_int),
        "obvi_factor_add_param_prior": ([vp, vp, C.c_int, dbl, dbl, dbl, u64p], C.c_int),
        "obvi_factor_remove": ([vp, C.c_uint64], C.c_int),
        "obvi_factor_remove_batch": ([vp, u64p, i64], C.c_int),
        "obvi_num_factors": ([vp], i64),
        "obvi_num_structure_builds": ([vp], i64),
        "obvi_residual_blocks": ([vp, u64p, i32p, i32p, i64, C.POINTER(i64)], C.c_int),
        "obvi_solver_options_init": ([C.POINTER(SolverOptions)], None),
        "obvi_solve": ([vp, C.POINTER(SolverOptions), C.POINTER(Summary), C.POINTER(IterationSummary), i32], C.c_int),
        "obvi_evaluate": ([vp, C.c_int, _d, _d, i64, C.POINTER(i64)], C.c_int),
        "obvi_evaluate_factor_type": ([vp, C.c_int, C.c_int, _d, _d, _d], C.c_int),
        "obvi_topk_outliers": ([vp, C.c_int, dbl, u64p, i64, C.POINTER(i64)], C.c_int),
        "obvi_evaluate_jacobian": ([vp, C.c_int, u64p, i64, C.POINTER(vp), i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), i32p, i32p, _d, _d], C.c_int),
        "obvi_object_covariances": ([vp, i64, C.POINTER(vp), C.POINTER(vp), _d], C.c_int),
        "obvi_profile_jacobian": ([vp, C.c_int, _d, C.POINTER(i64), C.POINTER(i64)], C.c_int),
        "obvi_debug_partition": ([vp, C.c_int, C.c_int, C.POINTER(i64)], C.c_int),
        "obvi_debug_structure_hash": ([vp, C.c_int, C.c_int, C.POINTER(C.c_uint64)], C.c_int),
        "obvi_debug_row_products": ([vp, C.c_int, C.c_int, C.POINTER(C.c_int64)], C.c_int),
        "obvi_comm_unique_id": ([vp], C.c_int),
        "obvi_comm_init": ([vp, vp, C.c_int, C.c_int], C.c_int),
        "obvi_comm_init_local": ([C.POINTER(vp), C.c_int], C.c_int),
        "obvi_comm_attach": ([vp, vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)  # AttributeError here means the library does not export what obvi_ba.h declares
        fn.argtypes = args
        fn.restype = res
    _LIB = L
    return L


EXPORTED_SYMBOLS = [
    "obvi_problem_create", "obvi_problem_destroy", "obvi_last_error", "obvi_version", "obvi_param_add",
    "obvi_param_add_array", "obvi_param_remove", "obvi_param_set_constant", "obvi_param_is_constant", "obvi_camera_add",
    "obvi_factor_add_reproj", "obvi_factor_add_reproj_batch", "obvi_factor_add_bbox", "obvi_factor_add_bbox_batch",
    "obvi_factor_add_shape_prior", "obvi_factor_add_ltm_prior", "obvi_factor_add_rel_pose", "obvi_factor_add_param_prior",
    "obvi_factor_remove", "obvi_factor_remove_batch", "obvi_num_factors", "obvi_num_structure_builds", "obvi_residual_blocks", "obvi_solver_options_init", "obvi_solve",
    "obvi_evaluate", "obvi_evaluate_factor_type", "obvi_topk_outliers", "obvi_evaluate_jacobian", "obvi_object_covariances", "obvi_profile_jacobian", "obvi_debug_partition", "obvi_debug_row_products", "obvi_debug_structure_hash", "obvi_comm_unique_id",
    "obvi_comm_init", "obvi_comm_init_local", "obvi_comm_attach",
]


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_d)


def _ptrs(arr, idx):
    """Host addresses of rows idx of a C-contiguous 2-D float64 array (the block identities)."""
    assert arr.dtype == np.float64 and arr.flags.c_contiguous
    return (arr.ctypes.data + np.asarray(idx, dtype=np.int64) * arr.strides[0]).astype(np.uint64)


class Problem:
    """Mirror of the ceres::Problem surface the reference drives (AddParameterBlock, AddResidualBlock via the
    factor factories, SetParameterBlockConstant/Variable, RemoveResidualBlock, Solve, Evaluate)."""

    def __init__(self, device=0):
        self._lib = lib()
        h = C.c_void_p()
        rc = self._lib.obvi_problem_create(int(device), C.byref(h))
        if rc != 0:
            raise ObviError(f"obvi_problem_create failed ({rc}): {self._lib.obvi_last_error(None).decode()}")
        self._h = h
        self._keep = []  # arrays whose memory the backend references (parameter blocks)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.obvi_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise ObviError(f"obvi error {rc}: {self._lib.obvi_last_error(self._h).decode()}")

    # ---- parameter blocks
    def add_parameter_array(self, arr):
        """Register every row of a (n, 3|6|7) float64 array as one parameter block."""
        assert arr.dtype == np.float64 and arr.flags.c_contiguous and arr.ndim == 2
        self._keep.append(arr)
        if len(arr):
            self._ck(self._lib.obvi_param_add_array(self._h, arr.ctypes.data, arr.shape[1], len(arr)))

    def add_parameter_block(self, block):
        self._keep.append(block)
        self._ck(self._lib.obvi_param_add(self._h, block.ctypes.data, block.size))

    def set_parameter_block_constant(self, block, constant=True):
        self._ck(self._lib.obvi_param_set_constant(self._h, block.ctypes.data, int(constant)))

    def set_parameter_block_variable(self, block):
        self.set_parameter_block_constant(block, False)

    def is_parameter_block_constant(self, block):
        out = C.c_int(0)
        self._ck(self._lib.obvi_param_is_constant(self._h, block.ctypes.data, C.byref(out)))
        return bool(out.value)

    def remove_parameter_block(self, block):
        self._ck(self._lib.obvi_param_remove(self._h, block.ctypes.data))

    def add_camera(self, intrinsics, R, t):
        _, pi = _f64(intrinsics); _, pr = _f64(R); _, pt = _f64(t)
        a, b, c = _f64(intrinsics), _f64(R), _f64(t)
        cid = C.c_int(-1)
        self._ck(self._lib.obvi_camera_add(self._h, a[1], b[1], c[1], C.byref(cid)))
        return cid.value

    # ---- residual blocks (names follow the reference factories)
    def add_reprojection(self, pose, point, cam, pixel, sigma, huber):
        a = _f64(pixel); fid = C.c_uint64(0)
        self._ck(self._lib.obvi_factor_add_reproj(self._h, pose.ctypes.data, point.ctypes.data, cam, a[1], sigma, huber, C.byref(fid)))
        return fid.value

    def add_reprojection_batch(self, poses, pose_idx, points, point_idx, cams, pixels, sigmas, huber):
        n = len(pose_idx)
        ids = np.zeros(n, dtype=np.uint64)
        if n == 0:
            return ids
        pp, qp = _ptrs(poses, pose_idx), _ptrs(points, point_idx)
        cams = np.ascontiguousarray(cams, dtype=np.int32)
        px, sg = _f64(pixels), _f64(sigmas)
        self._ck(self._lib.obvi_factor_add_reproj_batch(self._h, n, pp.ctypes.data, qp.ctypes.data,
                                                        cams.ctypes.data_as(C.POINTER(C.c_int32)), px[1], sg[1], float(huber),
                                                        ids.ctypes.data_as(C.POINTER(C.c_uint64))))
        return ids

    def add_bounding_box(self, ellipsoid, pose, cam, corners, cov, invalid_err, huber):
        a, b = _f64(corners), _f64(cov); fid = C.c_uint64(0)
        self._ck(self._lib.obvi_factor_add_bbox(self._h, ellipsoid.ctypes.data, pose.ctypes.data, cam, a[1], b[1], invalid_err, huber, C.byref(fid)))
        return fid.value

    def add_bounding_box_batch(self, objects, obj_idx, poses, pose_idx, cams, corners, covs, invalid_err, huber):
        n = len(obj_idx)
        ids = np.zeros(n, dtype=np.uint64)
        if n == 0:
            return ids
        op, pp = _ptrs(objects, obj_idx), _ptrs(poses, pose_idx)
        cams = np.ascontiguousarray(cams, dtype=np.int32)
        a, b = _f64(corners), _f64(covs)
        self._ck(self._lib.obvi_factor_add_bbox_batch(self._h, n, op.ctypes.data, pp.ctypes.data,
                                                      cams.ctypes.data_as(C.POINTER(C.c_int32)), a[1], b[1], float(invalid_err),
                                                      float(huber), ids.ctypes.data_as(C.POINTER(C.c_uint64))))
        return ids

    def add_shape_prior(self, ellipsoid, mean, cov, huber):
        a, b = _f64(mean), _f64(cov); fid = C.c_uint64(0)
        self._ck(self._lib.obvi_factor_add_shape_prior(self._h, ellipsoid.ctypes.data, a[1], b[1], huber, C.byref(fid)))
        return fid.value

    def add_ltm_prior(self, ellipsoid, mean, cov, huber):
        a, b = _f64(mean), _f64(cov); fid = C.c_uint64(0)
        self._ck(self._lib.obvi_factor_add_ltm_prior(self._h, ellipsoid.ctypes.data, a[1], b[1], huber, C.byref(fid)))
        return fid.value

    def add_relative_pose(self, pose1, pose2, t, R, cov, huber):
        a, b, c = _f64(t), _f64(R), _f64(cov); fid = C.c_uint64(0)
        self._ck(self._lib.obvi_factor_add_rel_pose(self._h, pose1.ctypes.data, pose2.ctypes.data, a[1], b[1], c[1], huber, C.byref(fid)))
        return fid.value

    def add_parameter_prior(self, block, idx, mean, std, huber=0.0):
        fid = C.c_uint64(0)
        self._ck(self._lib.obvi_factor_add_param_prior(self._h, block.ctypes.data, idx, mean, std, huber, C.byref(fid)))
        return fid.value

    def remove_residual_block(self, fid):
        self._ck(self._lib.obvi_factor_remove(self._h, C.c_uint64(int(fid))))

    def remove_residual_blocks(self, fids):
        ids = np.ascontiguousarray(fids, dtype=np.uint64)
        self._ck(self._lib.obvi_factor_remove_batch(self._h, ids.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int64(len(ids))))

    def num_residual_blocks(self):
        return int(self._lib.obvi_num_factors(self._h))

    def num_structure_builds(self):
        return int(self._lib.obvi_num_structure_builds(self._h))

    def residual_blocks(self):
        n = C.c_int64(0)
        self._ck(self._lib.obvi_residual_blocks(self._h, None, None, None, 0, C.byref(n)))
        ids = np.zeros(n.value, np.uint64); types = np.zeros(n.value, np.int32); sizes = np.zeros(n.value, np.int32)
        self._ck(self._lib.obvi_residual_blocks(self._h, ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                types.ctypes.data_as(C.POINTER(C.c_int32)),
                                                sizes.ctypes.data_as(C.POINTER(C.c_int32)), n.value, C.byref(n)))
        return ids, types, sizes

    # ---- solve / evaluate
    def solve(self, options=None, callback=None, **kw):
        """callback(dict) -> 0 continue | 1 abort | 2 terminate successfully: ceres::IterationCallback."""
        opt = options if options is not None else SolverOptions(**kw)
        if callback is not None:
            def tramp(_user, it_ptr):
                i = C.cast(it_ptr, C.POINTER(IterationSummary)).contents
                return int(callback(dict(iteration=i.iteration, cost=i.cost, cost_change=i.cost_change, step_norm=i.step_norm,
                                         successful=bool(i.step_is_successful), radius=i.trust_region_radius,
                                         gradient_max_norm=i.gradient_max_norm)) or 0)
            opt._cb = ITERATION_CALLBACK(tramp)     # keep the thunk alive for the duration of the call
            opt.iteration_callback = opt._cb
        summ = Summary()
        cap = opt.max_num_iterations + 2
        its = (IterationSummary * cap)()
        self._ck(self._lib.obvi_solve(self._h, C.byref(opt), C.byref(summ), its, cap))
        summ.iterations = [dict(iteration=i.iteration, cost=i.cost, cost_change=i.cost_change, step_norm=i.step_norm,
                                successful=bool(i.step_is_successful), valid=bool(i.step_is_valid), radius=i.trust_region_radius,
                                gradient_max_norm=i.gradient_max_norm, relative_decrease=i.relative_decrease,
                                linear_solver_iterations=i.linear_solver_iterations)
                           for i in its[:min(cap, summ.num_iterations)]]
        return summ

    def evaluate(self, apply_loss_function=True, residuals=True):
        cost = C.c_double(0); n = C.c_int64(0)
        if not residuals:
            self._ck(self._lib.obvi_evaluate(self._h, int(apply_loss_function), C.byref(cost), None, 0, None))
            return cost.value, None
        self._ck(self._lib.obvi_evaluate(self._h, int(apply_loss_function), C.byref(cost), None, 0, C.byref(n)))
        r = np.zeros(n.value)
        self._ck(self._lib.obvi_evaluate(self._h, int(apply_loss_function), C.byref(cost), r.ctypes.data_as(_d), n.value, C.byref(n)))
        return cost.value, r

    def evaluate_factor_type(self, ftype, n, apply_loss_function=False):
        """Residuals + Jacobians (Ceres layout) of the n live blocks of one type, in order of addition."""
        shapes = {FACTOR_REPROJECTION: (2, (2, 6), (2, 3)), FACTOR_BBOX: (4, (4, 7), (4, 6)), FACTOR_SHAPE_PRIOR: (3, (3, 7), None),
                  FACTOR_LTM_PRIOR: (7, (7, 7), None), FACTOR_REL_POSE: (6, (6, 6), (6, 6)), FACTOR_PARAM_PRIOR: (1, (1, 7), None)}
        k, s0, s1 = shapes[ftype]
        r = np.zeros((n, k)); J0 = np.zeros((n,) + s0); J1 = np.zeros((n,) + s1) if s1 else None
        self._ck(self._lib.obvi_evaluate_factor_type(self._h, ftype, int(apply_loss_function), r.ctypes.data_as(_d),
                                                     J0.ctypes.data_as(_d), J1.ctypes.data_as(_d) if J1 is not None else None))
        return r, J0, J1

    def topk_outliers(self, ftype, fraction):
        n = C.c_int64(0)
        cap = self.num_residual_blocks()
        ids = np.zeros(max(cap, 1), np.uint64)
        self._ck(self._lib.obvi_topk_outliers(self._h, ftype, float(fraction), ids.ctypes.data_as(C.POINTER(C.c_uint64)), cap, C.byref(n)))
        return ids[:n.value]

    def evaluate_jacobian(self, apply_loss_function=True, ids=None, blocks=None):
        """(rows, cols, values, shape, gradient) of the CRS Jacobian (Problem::Evaluate with a CRSMatrix)."""
        pid = None if ids is None else np.ascontiguousarray(ids, np.uint64)
        nid = 0 if ids is None else len(pid)
        pbl = None if blocks is None else (C.c_void_p * len(blocks))(*[b.ctypes.data for b in blocks])
        nbl = 0 if blocks is None else len(blocks)
        nr, nc, nz = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        idp = None if pid is None else pid.ctypes.data_as(C.POINTER(C.c_uint64))
        self._ck(self._lib.obvi_evaluate_jacobian(self._h, int(apply_loss_function), idp, nid, pbl, nbl, C.byref(nr), C.byref(nc), C.byref(nz), None, None, None, None))
        rows = np.zeros(nr.value + 1, np.int32); cols = np.zeros(max(nz.value, 1), np.int32); vals = np.zeros(max(nz.value, 1)); grad = np.zeros(max(nc.value, 1))
        i32 = C.POINTER(C.c_int32)
        self._ck(self._lib.obvi_evaluate_jacobian(self._h, int(apply_loss_function), idp, nid, pbl, nbl, C.byref(nr), C.byref(nc), C.byref(nz),
                                                  rows.ctypes.data_as(i32), cols.ctypes.data_as(i32), vals.ctypes.data_as(_d), grad.ctypes.data_as(_d)))
        return rows, cols[:nz.value], vals[:nz.value], (nr.value, nc.value), grad[:nc.value]

    def object_covariances(self, blocks_a, blocks_b):
        """7x7 blocks [a_i, b_i] of (J^T J)^-1 (ceres::Covariance on ellipsoid blocks); blocks are the registered arrays."""
        n = len(blocks_a)
        pa = (C.c_void_p * n)(*[b.ctypes.data for b in blocks_a]); pb = (C.c_void_p * n)(*[b.ctypes.data for b in blocks_b])
        out = np.zeros((n, 7, 7))
        self._ck(self._lib.obvi_object_covariances(self._h, n, pa, pb, out.ctypes.data_as(_d)))
        return out

    def profile_jacobian(self, reps=20):
        """(seconds per launch, algorithmic bytes per launch, observations) of the Jacobian-evaluation kernel."""
        sec = C.c_double(0); nb = C.c_int64(0); no = C.c_int64(0)
        self._ck(self._lib.obvi_profile_jacobian(self._h, int(reps), C.byref(sec), C.byref(nb), C.byref(no)))
        return sec.value, nb.value, no.value

    def debug_partition(self, rank, world):
        """Host-only: how the structure build shards the graph for (rank, world); see obvi_debug_partition."""
        st = (C.c_int64 * 12)()
        self._ck(self._lib.obvi_debug_partition(self._h, rank, world, st))
        keys = ["n_obs", "n_bbox", "n_unary", "n_rel", "nf", "n_upper", "points_here", "objects_here", "point_batches",
                "structure_checksum", "num_parameters_reduced", "num_residual_blocks_reduced"]
        return dict(zip(keys, list(st)))

    def debug_row_products(self, rank=0, world=1):
        """Host-only: the row-pair work lists of the point elimination; see obvi_debug_row_products."""
        st = (C.c_int64 * 8)()
        self._ck(self._lib.obvi_debug_row_products(self._h, rank, world, st))
        keys = ["entries", "items", "products", "dense_slots", "regular_points", "fallback_points", "entries_in_items", "longest_item"]
        return dict(zip(keys, list(st)))

    def debug_structure_hash(self, rank=0, world=1):
        """Host-only: hash of every array of the structure built for (rank, world); see obvi_debug_structure_hash."""
        h = C.c_uint64(0)
        self._ck(self._lib.obvi_debug_structure_hash(self._h, rank, world, C.byref(h)))
        return h.value

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        rc = lib().obvi_comm_unique_id(buf)
        if rc != 0:
            raise ObviError(f"obvi_comm_unique_id failed: {lib().obvi_last_error(None).decode()}")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._lib.obvi_comm_init(self._h, buf, rank, world))

    def comm_attach(self, src):
        """Use the communicator of problem `src` (same device)."""
        self._ck(self._lib.obvi_comm_attach(self._h, src._h))

    @staticmethod
    def comm_init_local(problems):
        """Join problems of this process into one sharded solve (problems[i] = rank i); drive each from its own thread."""
        arr = (C.c_void_p * len(problems))(*[p._h for p in problems])
        rc = lib().obvi_comm_init_local(arr, len(problems))
        if rc != 0:
            raise ObviError(f"obvi_comm_init_local failed ({rc})")


def problem_from_graph(g, device=0):
    """Build a Problem from a FactorGraph (synth.py).  The graph's arrays are the parameter blocks: solve()
    updates g.poses / g.points / g.objects in place, like Ceres does with the reference's pose-graph nodes."""
    p = Problem(device)
    p.add_parameter_array(g.poses); p.add_parameter_array(g.points); p.add_parameter_array(g.objects)
    cam_ids = [p.add_camera(c["intr"], c["R"], c["t"]) for c in g.cams]
    cam_map = np.asarray(cam_ids, dtype=np.int32)
    ids = {}
    rp = g.reproj
    ids["reproj"] = p.add_reprojection_batch(g.poses, rp["pose"], g.points, rp["point"], cam_map[rp["cam"]] if len(rp["cam"]) else rp["cam"],
                                             rp["px"], rp["sigma"], rp["huber"])
    bb = g.bbox
    ids["bbox"] = p.add_bounding_box_batch(g.objects, bb["obj"], g.poses, bb["pose"], cam_map[bb["cam"]] if len(bb["cam"]) else bb["cam"],
                                           bb["corners"], bb["cov"], bb["invalid_err"], bb["huber"])
    sh = g.shape
    ids["shape"] = np.array([p.add_shape_prior(g.objects[o], sh["mean"][i], sh["cov"][i], sh["huber"]) for i, o in enumerate(sh["obj"])], dtype=np.uint64)
    lt = g.ltm
    ids["ltm"] = np.array([p.add_ltm_prior(g.objects[o], lt["mean"][i], lt["cov"][i], lt["huber"]) for i, o in enumerate(lt["obj"])], dtype=np.uint64)
    rl = g.relpose
    ids["relpose"] = np.array([p.add_relative_pose(g.poses[a], g.poses[b], rl["t"][i], rl["Rm"][i], rl["cov"][i], rl["huber"])
                               for i, (a, b) in enumerate(zip(rl["p1"], rl["p2"]))], dtype=np.uint64)
    for k in np.nonzero(g.const_pose)[0]:
        p.set_parameter_block_constant(g.poses[k])
    for k in np.nonzero(g.const_point)[0]:
        p.set_parameter_block_constant(g.points[k])
    for k in np.nonzero(g.const_obj)[0]:
        p.set_parameter_block_constant(g.objects[k])
    p.factor_ids = ids
    return p
