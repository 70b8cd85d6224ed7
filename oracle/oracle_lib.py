"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/libba_oracle.so (the C++ CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_d = C.POINTER(C.c_double)
_i = C.POINTER(C.c_int32)
_u8 = C.POINTER(C.c_uint8)


class OracleGraph(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("P", C.c_int32), ("O", C.c_int32), ("C", C.c_int32),
        ("poses", _d), ("points", _d), ("objects", _d),
        ("const_pose", _u8), ("const_point", _u8), ("const_obj", _u8),
        ("cam_intr", _d), ("cam_R", _d), ("cam_t", _d),
        ("n_reproj", C.c_int64), ("rp_pose", _i), ("rp_point", _i), ("rp_cam", _i), ("rp_px", _d), ("rp_sigma", _d),
        ("rp_huber", C.c_double),
        ("n_bbox", C.c_int64), ("bb_obj", _i), ("bb_pose", _i), ("bb_cam", _i), ("bb_corners", _d), ("bb_cov", _d),
        ("bb_huber", C.c_double), ("bb_invalid", C.c_double),
        ("n_shape", C.c_int64), ("sh_obj", _i), ("sh_mean", _d), ("sh_cov", _d), ("sh_huber", C.c_double),
        ("n_ltm", C.c_int64), ("lt_obj", _i), ("lt_mean", _d), ("lt_cov", _d), ("lt_huber", C.c_double),
        ("n_rel", C.c_int64), ("rl_p1", _i), ("rl_p2", _i), ("rl_t", _d), ("rl_R", _d), ("rl_cov", _d),
        ("rl_huber", C.c_double),
    ]


class OracleOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("use_nonmonotonic_steps", C.c_int32),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("num_threads", C.c_int32)]


class OracleSummary(C.Structure):
    _fields_ = [("termination", C.c_int32), ("num_iterations", C.c_int32), ("lm_steps", C.c_int32),
                ("num_parameters_reduced", C.c_int32), ("num_threads_used", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("fixed_cost", C.c_double),
                ("total_time", C.c_double), ("linear_solver_time", C.c_double), ("jacobian_time", C.c_double),
                ("residual_time", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libba_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("ba_oracle.cpp", "ba_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "libba_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_solve.argtypes = [C.POINTER(OracleGraph), C.POINTER(OracleOptions), C.POINTER(OracleSummary), _d, C.c_int32]
        _LIB.oracle_solve.restype = C.c_int
        _LIB.oracle_evaluate.argtypes = [C.POINTER(OracleGraph), C.c_int, _d] + [_d] * 13
        _LIB.oracle_evaluate.restype = C.c_int
        _LIB.oracle_sqrt_information.argtypes = [_d, C.c_int, _d]
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def pack(g):
    """FactorGraph -> (OracleGraph, keepalive list).  poses/points/objects are referenced, not copied."""
    keep = []

    def f64(a):
        a = np.ascontiguousarray(a, dtype=np.float64); keep.append(a); return _p(a, _d)

    def i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32); keep.append(a); return _p(a, _i)

    def u8(a):
        a = np.ascontiguousarray(a, dtype=np.uint8); keep.append(a); return _p(a, _u8)

    og = OracleGraph()
    og.K, og.P, og.O, og.C = len(g.poses), len(g.points), len(g.objects), len(g.cams)
    for name in ("poses", "points", "objects"):
        arr = getattr(g, name)
        assert arr.dtype == np.float64 and arr.flags.c_contiguous
        setattr(og, name, _p(arr, _d))
    og.const_pose, og.const_point, og.const_obj = u8(g.const_pose), u8(g.const_point), u8(g.const_obj)
    og.cam_intr = f64(np.array([c["intr"] for c in g.cams]))
    og.cam_R = f64(np.array([c["R"] for c in g.cams]))
    og.cam_t = f64(np.array([c["t"] for c in g.cams]))
    rp = g.reproj
    og.n_reproj = len(rp["pose"]); og.rp_pose, og.rp_point, og.rp_cam = i32(rp["pose"]), i32(rp["point"]), i32(rp["cam"])
    og.rp_px, og.rp_sigma, og.rp_huber = f64(rp["px"]), f64(rp["sigma"]), float(rp["huber"])
    bb = g.bbox
    og.n_bbox = len(bb["obj"]); og.bb_obj, og.bb_pose, og.bb_cam = i32(bb["obj"]), i32(bb["pose"]), i32(bb["cam"])
    og.bb_corners, og.bb_cov, og.bb_huber, og.bb_invalid = f64(bb["corners"]), f64(bb["cov"]), float(bb["huber"]), float(bb["invalid_err"])
    sh = g.shape
    og.n_shape = len(sh["obj"]); og.sh_obj, og.sh_mean, og.sh_cov, og.sh_huber = i32(sh["obj"]), f64(sh["mean"]), f64(sh["cov"]), float(sh["huber"])
    lt = g.ltm
    og.n_ltm = len(lt["obj"]); og.lt_obj, og.lt_mean, og.lt_cov, og.lt_huber = i32(lt["obj"]), f64(lt["mean"]), f64(lt["cov"]), float(lt["huber"])
    rl = g.relpose
    og.n_rel = len(rl["p1"]); og.rl_p1, og.rl_p2 = i32(rl["p1"]), i32(rl["p2"])
    og.rl_t, og.rl_R, og.rl_cov, og.rl_huber = f64(rl["t"]), f64(rl["Rm"]), f64(rl["cov"]), float(rl["huber"])
    return og, keep


TERMINATION = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}


def solve(g, max_num_iterations=50, function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8,
          initial_radius=1e4, max_radius=1e16, use_nonmonotonic_steps=False, num_threads=0):
    """Runs the C++ oracle LM in place on g.poses / g.points / g.objects.  Returns a dict."""
    og, keep = pack(g)
    opt = OracleOptions(max_num_iterations, int(use_nonmonotonic_steps), function_tolerance, gradient_tolerance,
                        parameter_tolerance, initial_radius, max_radius, num_threads)
    summ = OracleSummary()
    log = np.zeros((max_num_iterations + 2, 7))
    rc = lib().oracle_solve(C.byref(og), C.byref(opt), C.byref(summ), _p(log, _d), log.shape[0])
    assert rc == 0
    its = [dict(iteration=int(r[0]), cost=r[1], cost_change=r[2], step_norm=r[3], successful=bool(r[4]), radius=r[5],
                gradient_max_norm=r[6]) for r in log[:summ.num_iterations]]
    return dict(iterations=its, termination=TERMINATION[summ.termination], lm_steps=summ.lm_steps,
                initial_cost=summ.initial_cost, final_cost=summ.final_cost, fixed_cost=summ.fixed_cost,
                num_parameters_reduced=summ.num_parameters_reduced, num_threads=summ.num_threads_used,
                total_time=summ.total_time, linear_solver_time=summ.linear_solver_time,
                jacobian_time=summ.jacobian_time, residual_time=summ.residual_time)


def evaluate(g, apply_loss=False):
    """Residuals and Jacobians per block (Ceres layout) from the C++ oracle."""
    og, keep = pack(g)
    n = dict(rp=og.n_reproj, bb=og.n_bbox, sh=og.n_shape, lt=og.n_ltm, rl=og.n_rel)
    out = dict(r_reproj=np.zeros((n["rp"], 2)), jp_reproj=np.zeros((n["rp"], 2, 6)), jl_reproj=np.zeros((n["rp"], 2, 3)),
               r_bbox=np.zeros((n["bb"], 4)), jo_bbox=np.zeros((n["bb"], 4, 7)), jp_bbox=np.zeros((n["bb"], 4, 6)),
               r_shape=np.zeros((n["sh"], 3)), j_shape=np.zeros((n["sh"], 3, 7)),
               r_ltm=np.zeros((n["lt"], 7)), j_ltm=np.zeros((n["lt"], 7, 7)),
               r_rel=np.zeros((n["rl"], 6)), j1_rel=np.zeros((n["rl"], 6, 6)), j2_rel=np.zeros((n["rl"], 6, 6)))
    cost = C.c_double(0)
    order = ["r_reproj", "jp_reproj", "jl_reproj", "r_bbox", "jo_bbox", "jp_bbox", "r_shape", "j_shape", "r_ltm", "j_ltm",
             "r_rel", "j1_rel", "j2_rel"]
    rc = lib().oracle_evaluate(C.byref(og), int(apply_loss), C.byref(cost), *[_p(out[k], _d) for k in order])
    assert rc == 0
    out["cost"] = cost.value
    return out


def sqrt_information(cov):
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    out = np.zeros_like(cov)
    lib().oracle_sqrt_information(_p(cov, _d), cov.shape[0], _p(out, _d))
    return out
