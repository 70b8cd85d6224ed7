"""TEST INFRASTRUCTURE ONLY -- NumPy oracle (a) for the ObVi-SLAM bundle-adjustment hot path.

PARITY UNPINNED: the reference holds no golden vectors for this path and its solver
arithmetic lives in Ceres/SuiteSparse, which are absent here.
(Only the projection model of the reprojection factor is pinned by reference material:
the simulated sequences under /root/reference/data, see tests/test_vslam_dataset.py.)
This file restates the reference's *residual formulas* (what Ceres autodiff evaluates)
in float64 NumPy, differentiates them by the complex-step method (exact to rounding),
and restates Ceres' Levenberg-Marquardt semantics with a *dense* normal-equation solve.
It is structurally different from oracle/ba_oracle.cpp (dual numbers + Schur + sparse
Cholesky), so that the two can check each other.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Reference formulas followed (paths relative to /root/reference):
  reprojection   include/refactoring/factors/reprojection_cost_functor.h:56-93,
                 include/refactoring/types/vslam_math_util.h:347-394,
                 src/refactoring/factors/reprojection_cost_functor.cpp:5-17
  bounding box   include/refactoring/factors/bounding_box_factor.h:68-136,
                 include/refactoring/types/ellipsoid_utils.h:159-273,
                 src/refactoring/factors/bounding_box_factor.cpp:7-40
  shape prior    include/refactoring/factors/shape_prior_factor.h:46-61
  relative pose  include/refactoring/factors/relative_pose_factor.h:32-61,
                 include/refactoring/types/vslam_math_util.h:121-141
  LTM prior      include/refactoring/factors/independent_object_map_factor.h:21-33
  param prior    include/refactoring/factors/parameter_prior.h:27-34
  odom cov       include/refactoring/factors/relative_pose_factor_utils.h:17-36
"""
from __future__ import annotations

import numpy as np

K_SMALL_ANGLE = 1e-8  # vslam_math_util.h:17
K_DIM_REG = float(np.float32(1e-3))  # ellipsoid_utils.h:22 -- a float constant promoted to double


# ----------------------------------------------------------------------------- helpers
def _norm3(v):
    """Analytic (complex-step safe) Euclidean norm."""
    return np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def rot_from_angle_axis(angle, axis):
    """Eigen::AngleAxis::toRotationMatrix (Rodrigues).  angle scalar, axis 3-vector."""
    c = np.cos(angle)
    s = np.sin(angle)
    t = 1.0 - c
    x, y, z = axis
    return np.array(
        [
            [t * x * x + c, t * x * y - s * z, t * x * z + s * y],
            [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
            [t * x * z - s * y, t * y * z + s * x, t * z * z + c],
        ]
    )


def inv_rot_functor(omega):
    """R(omega)^T as built by the reprojection / bbox functors: AngleAxis(-|w|, w/|w|) when
    |w| > 1e-8, else the *constant* identity (vslam_math_util.h:361-369)."""
    ang = _norm3(omega)
    if np.real(ang) > K_SMALL_ANGLE:
        return rot_from_angle_axis(-ang, omega / ang)
    return np.eye(3, dtype=omega.dtype)


def rot_pose_array(omega):
    """PoseArrayToAffine rotation (vslam_math_util.h:121-141): identity iff |w| < 1e-8."""
    ang = _norm3(omega)
    if np.real(ang) < K_SMALL_ANGLE:
        return np.eye(3, dtype=omega.dtype)
    return rot_from_angle_axis(ang, omega / ang)


def sqrt_inv_spd(cov):
    """Principal square root of the inverse: Eigen's cov.inverse().sqrt() for SPD input."""
    cov = np.asarray(cov, dtype=np.float64)
    w, v = np.linalg.eigh(0.5 * (cov + cov.T))
    return (v * (1.0 / np.sqrt(w))) @ v.T


def _atan2_cs(y, x):
    """atan2 with a first-order complex extension (what the complex-step method needs)."""
    yr, xr = np.real(y), np.real(x)
    re = np.arctan2(yr, xr)
    if np.iscomplexobj(y) or np.iscomplexobj(x):
        im = (xr * np.imag(y) - yr * np.imag(x)) / (xr * xr + yr * yr)
        return re + 1j * im
    return re


def quat_from_rot(m):
    """Eigen Quaternion(Matrix3) -- Shepperd's method, same branch order.  Returns (w, x, y, z)."""
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if np.real(t) > 0:
        t = np.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        return w, (m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t
    i = 0
    if np.real(m[1, 1]) > np.real(m[0, 0]):
        i = 1
    if np.real(m[2, 2]) > np.real(m[i, i]):
        i = 2
    j = (i + 1) % 3
    k = (j + 1) % 3
    t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
    q = [None, None, None]
    q[i] = 0.5 * t
    t = 0.5 / t
    w = (m[k, j] - m[j, k]) * t
    q[j] = (m[j, i] + m[i, j]) * t
    q[k] = (m[k, i] + m[i, k]) * t
    return w, q[0], q[1], q[2]


def angle_axis_vec_from_rot(m):
    """angle * axis of Eigen::AngleAxis(Matrix3) (via quaternion; relative_pose_factor.h:53-55)."""
    w, x, y, z = quat_from_rot(m)
    n = np.sqrt(x * x + y * y + z * z)
    if np.real(n) != 0.0:
        wr = np.real(w)
        absw = w if wr >= 0 else -w
        ang = 2.0 * _atan2_cs(n, absw)
        if wr < 0:
            n = -n
        return np.array([ang * x / n, ang * y / n, ang * z / n])
    return np.zeros(3, dtype=np.result_type(w, np.float64))  # angle 0 * axis (1,0,0)


# ----------------------------------------------------------------------------- residuals
def reproj_residual(pose, point, px, intr, extr_R, extr_t, sigma):
    """A.1.  intr = (fx, fy, cx, cy); extrinsics = camera pose in the robot frame."""
    fx, fy, cx, cy = intr
    Rt = inv_rot_functor(pose[3:6])
    p_robot = Rt @ point - Rt @ pose[0:3]
    p_cam = extr_R.T @ p_robot - extr_R.T @ extr_t
    u = p_cam[0] / p_cam[2]
    v = p_cam[1] / p_cam[2]
    return np.array(
        [(fx / sigma) * (u - (px[0] - cx) / fx), (fy / sigma) * (v - (px[1] - cy) / fy)]
    )


def bbox_corners_rectified(ell, pose, extr_R, extr_t):
    """getCornerLocationsVectorRectified; returns None in the invalid case."""
    Rt = inv_rot_functor(pose[3:6])
    R_cw = extr_R.T @ Rt
    t_cw = extr_R.T @ (-(Rt @ pose[0:3])) - extr_R.T @ extr_t
    d = np.array(
        [
            (ell[4] / 2.0) ** 2 + K_DIM_REG,
            (ell[5] / 2.0) ** 2 + K_DIM_REG,
            (ell[6] / 2.0) ** 2 + K_DIM_REG,
            -1.0,
        ]
    )
    # yaw-only rotation through a quaternion, as Eigen does (ellipsoid_utils.h:218-227)
    hw, hz = np.cos(ell[3] / 2.0), np.sin(ell[3] / 2.0)
    Rz = np.array(
        [
            [1.0 - 2.0 * hz * hz, -2.0 * hw * hz, 0.0 * hw],
            [2.0 * hw * hz, 1.0 - 2.0 * hz * hz, 0.0 * hw],
            [0.0 * hw, 0.0 * hw, 1.0 + 0.0 * hw],
        ]
    )
    P = np.concatenate([R_cw @ Rz, (R_cw @ ell[0:3] + t_cw)[:, None]], axis=1)  # 3x4
    Q = (P * d[None, :]) @ P.T
    xin = Q[0, 2] ** 2 - Q[0, 0] * Q[2, 2]
    yin = Q[1, 2] ** 2 - Q[1, 1] * Q[2, 2]
    if np.real(xin) <= 0 or np.real(yin) <= 0:
        return None
    xs, ys = np.sqrt(xin), np.sqrt(yin)
    return np.array([Q[0, 2] + xs, Q[0, 2] - xs, Q[1, 2] + ys, Q[1, 2] - ys]) / Q[2, 2]


def bbox_residual(ell, pose, corners_px, cov4, intr, extr_R, extr_t, invalid_err):
    """A.2.  corners_px = (xmin, xmax, ymin, ymax)."""
    fx, fy, cx, cy = intr
    A = sqrt_inv_spd(cov4) @ np.diag([fx, fx, fy, fy])
    b = np.array(
        [
            (corners_px[0] - cx) / fx,
            (corners_px[1] - cx) / fx,
            (corners_px[2] - cy) / fy,
            (corners_px[3] - cy) / fy,
        ]
    )
    c = bbox_corners_rectified(ell, pose, extr_R, extr_t)
    if c is None:
        return np.full(4, invalid_err, dtype=np.result_type(ell.dtype, pose.dtype))
    return A @ (c - b)


def shape_residual(ell, mean, cov3):
    return sqrt_inv_spd(cov3) @ (ell[4:7] - mean)


def ltm_residual(ell, mean7, cov7):
    return sqrt_inv_spd(cov7) @ (ell - mean7)


def param_prior_residual(block, idx, mean, std):
    return np.array([(block[idx] - mean) / std])


def relpose_residual(pose1, pose2, meas_t, meas_R, cov6):
    """A.4.  meas_R is the measured rotation matrix; its general inverse is used."""
    R1 = rot_pose_array(pose1[3:6])
    R2 = rot_pose_array(pose2[3:6])
    t12 = R1.T @ (pose2[0:3] - pose1[0:3])
    R12 = R1.T @ R2
    Rerr = R12 @ np.linalg.inv(np.asarray(meas_R, dtype=np.float64))
    un = np.concatenate([t12 - meas_t, angle_axis_vec_from_rot(Rerr)])
    return sqrt_inv_spd(cov6) @ un


def generate_odom_cov(rel_t, rel_angle, rel_axis, k_tt, k_tr, k_rt, k_rr):
    """generateOdomCov (relative_pose_factor_utils.h:17-36).  k_tr = transl_error_mult_for_rot_error."""
    sd = np.empty(6)
    sd[0:3] = np.abs(rel_t) * k_tt + abs(rel_angle) * k_rt
    sd[3:6] = np.abs(rel_axis * rel_angle) * k_rr + np.linalg.norm(rel_t) * k_tr
    var = np.maximum(sd**2, (1e-3) ** 2)
    return np.diag(var)


# ----------------------------------------------------------------------------- derivatives
def complex_step_jacobian(fun, blocks, h=1e-30):
    """Jacobians of fun(*blocks) w.r.t. each block by the complex-step method."""
    r0 = np.real(fun(*[np.asarray(b, dtype=np.float64) for b in blocks]))
    jacs = []
    for bi, b in enumerate(blocks):
        J = np.zeros((r0.size, len(b)))
        for k in range(len(b)):
            args = [np.asarray(x, dtype=np.complex128).copy() for x in blocks]
            args[bi][k] += 1j * h
            J[:, k] = np.imag(fun(*args)) / h
        jacs.append(J)
    return r0, jacs


# ----------------------------------------------------------------------------- Huber
def huber_rho(s, a):
    """Ceres HuberLoss::Evaluate -> (rho, rho', rho'')."""
    b = a * a
    if np.isfinite(b) and s > b:
        r = np.sqrt(s)
        rho1 = max(np.finfo(np.float64).tiny, a / r)
        return 2.0 * a * r - b, rho1, -rho1 / (2.0 * s)
    return s, 1.0, 0.0


# ----------------------------------------------------------------------------- graph glue
class Graph:
    """A small factor graph in plain NumPy (the interchange format used by tests).

    Attributes (all float64 / int64 NumPy arrays):
      poses (K,6), points (P,3), objects (O,7)
      cams: list of dict(intr=(fx,fy,cx,cy), R=(3,3), t=(3,))
      reproj: dict(pose, point, cam (N,), px (N,2), sigma (N,), huber)
      bbox:   dict(obj, pose, cam (M,), corners (M,4), cov (M,4,4), huber, invalid_err)
      shape:  dict(obj (S,), mean (S,3), cov (S,3,3), huber)
      relpose: dict(p1, p2 (R,), t (R,3), Rm (R,3,3), cov (R,6,6), huber)
      ltm:    dict(obj (L,), mean (L,7), cov (L,7,7), huber)
      const_pose / const_point / const_obj: boolean masks
    """


def residual_blocks(g):
    """Yield (kind, huber, fun, [(block_kind, index), ...]) for every residual block, in the order
    reproj, bbox, shape, ltm, relpose (the order the CUDA path and the C++ oracle also use)."""
    out = []
    rp = g.reproj
    for n in range(len(rp["pose"])):
        cam = g.cams[int(rp["cam"][n])]
        f = (lambda px, cam, sig: lambda pose, point: reproj_residual(
            pose, point, px, cam["intr"], cam["R"], cam["t"], sig))(rp["px"][n], cam, rp["sigma"][n])
        out.append(("reproj", rp["huber"], f, [("pose", int(rp["pose"][n])), ("point", int(rp["point"][n]))]))
    bb = g.bbox
    for n in range(len(bb["obj"])):
        cam = g.cams[int(bb["cam"][n])]
        f = (lambda c, cov, cam: lambda ell, pose: bbox_residual(
            ell, pose, c, cov, cam["intr"], cam["R"], cam["t"], bb["invalid_err"]))(bb["corners"][n], bb["cov"][n], cam)
        out.append(("bbox", bb["huber"], f, [("obj", int(bb["obj"][n])), ("pose", int(bb["pose"][n]))]))
    sh = g.shape
    for n in range(len(sh["obj"])):
        f = (lambda m, c: lambda ell: shape_residual(ell, m, c))(sh["mean"][n], sh["cov"][n])
        out.append(("shape", sh["huber"], f, [("obj", int(sh["obj"][n]))]))
    lt = g.ltm
    for n in range(len(lt["obj"])):
        f = (lambda m, c: lambda ell: ltm_residual(ell, m, c))(lt["mean"][n], lt["cov"][n])
        out.append(("ltm", lt["huber"], f, [("obj", int(lt["obj"][n]))]))
    rl = g.relpose
    for n in range(len(rl["p1"])):
        f = (lambda t, R, c: lambda a, b: relpose_residual(a, b, t, R, c))(rl["t"][n], rl["Rm"][n], rl["cov"][n])
        out.append(("relpose", rl["huber"], f, [("pose", int(rl["p1"][n])), ("pose", int(rl["p2"][n]))]))
    # ParameterPrior (parameter_prior.h:27-34), used by the long-term-map rank repair: optional `g.prior` =
    # dict(kind ("pose" | "point" | "obj"), index, idx, mean, std (N,)); added with no loss function
    pr = getattr(g, "prior", None)
    if pr is not None:
        for n in range(len(pr["index"])):
            f = (lambda i, m, sd: lambda block: param_prior_residual(block, i, m, sd))(int(pr["idx"][n]), float(pr["mean"][n]), float(pr["std"][n]))
            out.append(("prior", np.inf, f, [(pr["kind"][n], int(pr["index"][n]))]))
    return out


def _layout(g):
    """Column offsets of the variable blocks: poses, then points, then objects."""
    off = {}
    n = 0
    for kind, arr, const in (("pose", g.poses, g.const_pose), ("point", g.points, g.const_point),
                             ("obj", g.objects, g.const_obj)):
        for i in range(arr.shape[0]):
            if not const[i]:
                off[(kind, i)] = n
                n += arr.shape[1]
    return off, n


def _get(g, kind, i, x=None, off=None):
    arr = {"pose": g.poses, "point": g.points, "obj": g.objects}[kind]
    if x is not None and (kind, i) in off:
        o = off[(kind, i)]
        return x[o:o + arr.shape[1]]
    return arr[i]


def evaluate(g, x=None, off=None, want_jac=True, apply_loss=True):
    """Cost, stacked (corrected) residuals and dense Jacobian over the variable blocks."""
    blocks = residual_blocks(g)
    if off is None:
        off, _ = _layout(g)
    ncol = (max(off.values()) + 7) if off else 0
    rows, Js = [], []
    cost = 0.0
    for kind, a, fun, refs in blocks:
        vals = [np.array(_get(g, k, i, x, off), dtype=np.float64) for k, i in refs]
        if want_jac:
            r, jacs = complex_step_jacobian(fun, vals)
        else:
            r, jacs = np.real(fun(*vals)), None
        s = float(r @ r)
        if apply_loss:
            rho, rho1, _ = huber_rho(s, a)
            cost += 0.5 * rho
            sc = np.sqrt(rho1)  # rho'' <= 0 for Huber: corrector reduces to sqrt(rho') scaling
        else:
            cost += 0.5 * s
            sc = 1.0
        rows.append(sc * r)
        if want_jac:
            Jrow = np.zeros((r.size, ncol))
            for (k, i), J in zip(refs, jacs):
                if (k, i) in off:
                    Jrow[:, off[(k, i)]:off[(k, i)] + J.shape[1]] = sc * J
            Js.append(Jrow)
    r = np.concatenate(rows) if rows else np.zeros(0)
    J = np.concatenate(Js, axis=0) if want_jac and Js else None
    return cost, r, J


def solve_lm_dense(g, max_num_iterations=50, function_tolerance=1e-6, gradient_tolerance=1e-10,
                   parameter_tolerance=1e-8, initial_radius=1e4, max_radius=1e16,
                   use_nonmonotonic_steps=False, write_back=True):
    """Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy semantics (SURVEY.md Appendix B --
    external knowledge of upstream Ceres) with a dense solve of (J~^T J~ + D^2) y = J~^T r.
    Returns a dict with the per-iteration log (cost, cost_change, step_norm, successful, radius)."""
    off, n = _layout(g)
    # drop variable blocks no residual touches (Ceres removes them from the reduced program)
    touched = set()
    for _, _, _, refs in residual_blocks(g):
        touched.update(r for r in refs if r in off)
    off2, n = {}, 0
    for key in off:  # insertion order preserved
        if key in touched:
            off2[key] = n
            n += {"pose": 6, "point": 3, "obj": 7}[key[0]]
    off = off2
    x = np.zeros(n)
    for (k, i), o in off.items():
        b = _get(g, k, i)
        x[o:o + len(b)] = b

    def ev(xv, jac):
        c, r, J = evaluate(g, xv, off, want_jac=jac)
        return c, r, (J[:, :n] if J is not None else None)

    cost, r, J = ev(x, True)
    grad = J.T @ r
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0)))
    its = [dict(iteration=0, cost=cost, cost_change=0.0, step_norm=0.0, successful=False, radius=initial_radius,
                gradient_max_norm=float(np.abs(grad).max()) if n else 0.0)]
    # "iterations" mirrors Solver::Summary::iterations: the iteration that trips the parameter /
    # function tolerance returns before its summary is pushed, so it is counted only in lm_steps.
    out = dict(iterations=its, termination="NO_CONVERGENCE", initial_cost=cost, lm_steps=0)
    x_min, min_cost = x.copy(), cost
    if n == 0 or np.abs(grad).max() <= gradient_tolerance:
        out["termination"] = "CONVERGENCE"
        out["final_cost"] = cost
        return out
    radius, decrease = initial_radius, 2.0
    max_nonmono = 5 if use_nonmonotonic_steps else 0
    minimum = current = reference = candidate = cost
    acc_ref = acc_cand = 0.0
    n_nonmono = 0
    reuse_diag, diag = False, None
    n_invalid = 0
    it = 0
    while True:
        if it >= max_num_iterations:
            break
        it += 1
        out["lm_steps"] += 1
        Js = J * scale[None, :]
        if not reuse_diag:
            diag = np.clip((Js * Js).sum(axis=0), 1e-6, 1e32)
        D2 = diag / radius
        H = Js.T @ Js + np.diag(D2)
        try:
            y = np.linalg.solve(H, Js.T @ r)
            ok = np.all(np.isfinite(y))
        except np.linalg.LinAlgError:
            ok = False
        rec = dict(iteration=it, radius=radius)
        if ok:
            delta_s = -y
            Jd = Js @ delta_s
            model_change = -float(Jd @ (r + 0.5 * Jd))
            ok = model_change > 0
        if not ok:
            n_invalid += 1
            rec.update(cost=cost, cost_change=0.0, step_norm=0.0, successful=False)
            its.append(rec)
            if n_invalid >= 5:
                out["termination"] = "FAILURE"
                break
            radius /= decrease  # LevenbergMarquardtStrategy::StepIsInvalid () = StepRejected (0)
            decrease *= 2.0
            reuse_diag = True
            continue
        n_invalid = 0
        delta = delta_s * scale
        x_cand = x + delta
        cost_cand, _, _ = ev(x_cand, False)
        step_norm = float(np.linalg.norm(delta))
        rec.update(step_norm=step_norm)
        if step_norm <= parameter_tolerance * (np.linalg.norm(x) + parameter_tolerance):
            out["termination"] = "CONVERGENCE"
            break
        cost_change = cost - cost_cand
        if abs(cost_change) <= function_tolerance * cost:
            out["termination"] = "CONVERGENCE"
            break
        rho_now = (current - cost_cand) / model_change
        rho_hist = (reference - cost_cand) / (acc_ref + model_change)
        rho = max(rho_now, rho_hist)
        if rho > 1e-3:
            x = x_cand
            cost, r, J = ev(x, True)
            grad = J.T @ r
            radius = min(max_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease = 2.0
            reuse_diag = False
            current = cost_cand
            acc_ref += model_change
            acc_cand += model_change
            if current < minimum:
                minimum = current
                n_nonmono = 0
                candidate = current
                acc_cand = 0.0
            else:
                n_nonmono += 1
                if current > candidate:
                    candidate = current
                    acc_cand = 0.0
            if n_nonmono == max_nonmono:
                reference = candidate
                acc_ref = acc_cand
            rec.update(cost=cost, cost_change=cost_change, successful=True,
                       gradient_max_norm=float(np.abs(grad).max()))
            its.append(rec)
            if cost < min_cost:
                min_cost, x_min = cost, x.copy()
            if np.abs(grad).max() <= gradient_tolerance:
                out["termination"] = "CONVERGENCE"
                break
        else:
            radius /= decrease
            decrease *= 2.0
            reuse_diag = True
            rec.update(cost=cost_cand, cost_change=cost_change, successful=False)
            its.append(rec)
        if radius <= 1e-32:
            out["termination"] = "CONVERGENCE"
            break
    out["final_cost"] = min_cost
    if write_back:
        for (k, i), o in off.items():
            b = _get(g, k, i)
            b[:] = x_min[o:o + len(b)]
    return out


def covariance_blocks(g, pairs):
    """ceres::Covariance on ellipsoid blocks (long_term_object_map_extraction.cpp:362-440): blocks [(a, b)] of (J^T J)^-1
    with the loss-corrected Jacobian over the variable blocks; dense inverse (small graphs only)."""
    off, n = _layout(g)
    _, _, J = evaluate(g, apply_loss=True)
    J = J[:, :n]
    used = np.abs(J).sum(axis=0) > 0          # blocks without residuals are not part of the reduced program
    cov = np.zeros((n, n))
    cov[np.ix_(used, used)] = np.linalg.inv(J[:, used].T @ J[:, used])
    out = np.zeros((len(pairs), 7, 7))
    for i, (a, b) in enumerate(pairs):
        if ("obj", a) in off and ("obj", b) in off:
            oa, ob_ = off[("obj", a)], off[("obj", b)]
            out[i] = cov[oa:oa + 7, ob_:ob_ + 7]
    return out
