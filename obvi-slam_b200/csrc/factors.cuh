// Residuals and hand-derived analytic Jacobians of the ObVi-SLAM factors (float64).
//
// Each function states the reference functor it replaces; the reference evaluates these through
// ceres::AutoDiffCostFunction, here the derivatives are closed-form (chain rule through the same
// formulas, so the constant small-angle / invalid-ellipse branches keep autodiff's zero derivatives).
// All functions are __host__ __device__ so that tests/hostcheck can run the very same arithmetic on
// the CPU against the oracle; the product only ever calls them from CUDA kernels.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define OBVI_HD __host__ __device__ __forceinline__
#define OBVI_ALIGN16 __align__(16)
#else
#define OBVI_HD inline
#define OBVI_ALIGN16 alignas(16)
#endif

namespace obvi {

constexpr double kSmallAngle = 1e-8;  // vslam_math_util.h:17
// ellipsoid_utils.h:22 -- `const float kDimensionRegularizationConstant = 1e-3` promoted to double
constexpr double kDimReg = 0.001000000047497451305389404296875;

// Per-(pose, camera) quantities shared by every observation made from that pose with that camera.
//   X_cam = Rcw X_world + tcw,   d X_cam / d omega_k = M[k] X_world + m[k]   (zero in the small-angle branch)
struct PoseCam {
  double Rcw[9];
  double tcw[3];
  double M[27];
  double m[9];
};  // 48 doubles = 384 B

// R(omega) = Eigen::AngleAxis(|w|, w/|w|).toRotationMatrix() and dR/d omega_k (k = 0..2), row-major.
// `identity` selects the caller's constant-identity branch (derivative zero, as autodiff sees it).
OBVI_HD void rodrigues_with_derivs(const double* w, bool identity, double* R, double* dR /*27*/) {
  if (identity) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 27; i++) dR[i] = 0.0;
    return;
  }
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double a[3] = {w[0] / th, w[1] / th, w[2] / th};
  const double s = sin(th), c = cos(th), t = 1.0 - c;
  R[0] = t * a[0] * a[0] + c;        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
  R[3] = t * a[0] * a[1] + s * a[2]; R[4] = t * a[1] * a[1] + c;        R[5] = t * a[1] * a[2] - s * a[0];
  R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = t * a[2] * a[2] + c;
  const double ith = 1.0 / th;
  for (int k = 0; k < 3; k++) {
    // d theta = a_k, d a = (e_k - a a_k) / theta
    const double dth = a[k];
    double da[3] = {-a[0] * a[k] * ith, -a[1] * a[k] * ith, -a[2] * a[k] * ith};
    da[k] += ith;
    const double ds = c * dth, dc = -s * dth, dt = s * dth;
    double* D = dR + 9 * k;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        D[3 * i + j] = dt * a[i] * a[j] + t * (da[i] * a[j] + a[i] * da[j]) + ((i == j) ? dc : 0.0);
    // + d(s [a]x)
    const double v[3] = {ds * a[0] + s * da[0], ds * a[1] + s * da[1], ds * a[2] + s * da[2]};
    D[1] -= v[2]; D[2] += v[1]; D[3] += v[2]; D[5] -= v[0]; D[6] -= v[1]; D[7] += v[0];
  }
}

// Build the PoseCam entry from a raw pose (t, omega) and the inverse extrinsics (Re_inv = R_e^T,
// te_inv = -R_e^T t_e).  Branch: |omega| > 1e-8 else constant identity (vslam_math_util.h:361-369,
// ellipsoid_utils.h:176-184).  With jac == false only Rcw / tcw are produced.
OBVI_HD void make_pose_cam(const double* pose, const double* Re_inv, const double* te_inv, bool jac, PoseCam* out) {
  const double* w = pose + 3;
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double R[9], dR[27];
  rodrigues_with_derivs(w, !(th > kSmallAngle), R, dR);
  // Rcw = Re_inv R^T
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      out->Rcw[3 * i + j] = Re_inv[3 * i] * R[3 * j] + Re_inv[3 * i + 1] * R[3 * j + 1] + Re_inv[3 * i + 2] * R[3 * j + 2];
  for (int i = 0; i < 3; i++)
    out->tcw[i] = te_inv[i] - (out->Rcw[3 * i] * pose[0] + out->Rcw[3 * i + 1] * pose[1] + out->Rcw[3 * i + 2] * pose[2]);
  if (!jac) return;
  for (int k = 0; k < 3; k++) {
    const double* D = dR + 9 * k;
    double* Mk = out->M + 9 * k;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        Mk[3 * i + j] = Re_inv[3 * i] * D[3 * j] + Re_inv[3 * i + 1] * D[3 * j + 1] + Re_inv[3 * i + 2] * D[3 * j + 2];
    for (int i = 0; i < 3; i++)
      out->m[3 * k + i] = -(Mk[3 * i] * pose[0] + Mk[3 * i + 1] * pose[1] + Mk[3 * i + 2] * pose[2]);
  }
}

// Ceres HuberLoss(a) + Corrector for s = |r|^2: returns 0.5 * rho(s) and the residual / Jacobian
// scale sqrt(rho').  rho'' <= 0 for Huber, so Ceres' corrector reduces to that scaling.
OBVI_HD double huber(double a, double s, double* scale) {
  const double b = a * a;
  if (s <= b) { *scale = 1.0; return 0.5 * s; }
  const double rt = sqrt(s);
  double rho1 = a / rt;
  if (rho1 < 2.2250738585072014e-308) rho1 = 2.2250738585072014e-308;
  *scale = sqrt(rho1);
  return 0.5 * (2.0 * a * rt - b);
}

// ---------------------------------------------------------------------------------------------
// Reprojection residual -- ReprojectionCostFunctor::runOperator (reprojection_cost_functor.h:56-93)
// via getProjectedPixelLocationRectified (vslam_math_util.h:347-394).  (ur, vr) = rectified feature,
// (mx, my) = (fx, fy) / sigma (reprojection_cost_functor.cpp:10-16).  No depth guard, as the reference.
// Jp is 2x6 row-major (translation columns 0-2, rotation 3-5), Jl is 2x3.
OBVI_HD void reproj_residual(const PoseCam& pc, const double* X, double ur, double vr, double mx, double my, double* r) {
  const double x = pc.Rcw[0] * X[0] + pc.Rcw[1] * X[1] + pc.Rcw[2] * X[2] + pc.tcw[0];
  const double y = pc.Rcw[3] * X[0] + pc.Rcw[4] * X[1] + pc.Rcw[5] * X[2] + pc.tcw[1];
  const double z = pc.Rcw[6] * X[0] + pc.Rcw[7] * X[1] + pc.Rcw[8] * X[2] + pc.tcw[2];
  const double iz = 1.0 / z;  // same expression as in reproj_residual_jacobian: bit-identical residuals
  r[0] = mx * (x * iz - ur);
  r[1] = my * (y * iz - vr);
}

OBVI_HD void reproj_residual_jacobian(const PoseCam& pc, const double* X, double ur, double vr, double mx, double my,
                                      double* r, double* Jp, double* Jl) {
  const double x = pc.Rcw[0] * X[0] + pc.Rcw[1] * X[1] + pc.Rcw[2] * X[2] + pc.tcw[0];
  const double y = pc.Rcw[3] * X[0] + pc.Rcw[4] * X[1] + pc.Rcw[5] * X[2] + pc.tcw[1];
  const double z = pc.Rcw[6] * X[0] + pc.Rcw[7] * X[1] + pc.Rcw[8] * X[2] + pc.tcw[2];
  const double iz = 1.0 / z, u = x * iz, v = y * iz;
  r[0] = mx * (u - ur);
  r[1] = my * (v - vr);
  const double a0 = mx * iz, a2 = -mx * u * iz;  // d r0 / d Xc = (a0, 0, a2)
  const double b1 = my * iz, b2 = -my * v * iz;  // d r1 / d Xc = (0, b1, b2)
  for (int j = 0; j < 3; j++) {
    Jl[j] = a0 * pc.Rcw[j] + a2 * pc.Rcw[6 + j];
    Jl[3 + j] = b1 * pc.Rcw[3 + j] + b2 * pc.Rcw[6 + j];
    Jp[j] = -Jl[j];
    Jp[6 + j] = -Jl[3 + j];
  }
  for (int k = 0; k < 3; k++) {
    const double* Mk = pc.M + 9 * k;
    const double dx = Mk[0] * X[0] + Mk[1] * X[1] + Mk[2] * X[2] + pc.m[3 * k];
    const double dy = Mk[3] * X[0] + Mk[4] * X[1] + Mk[5] * X[2] + pc.m[3 * k + 1];
    const double dz = Mk[6] * X[0] + Mk[7] * X[1] + Mk[8] * X[2] + pc.m[3 * k + 2];
    Jp[3 + k] = a0 * dx + a2 * dz;
    Jp[9 + k] = b1 * dy + b2 * dz;
  }
}

// ---------------------------------------------------------------------------------------------
// Bounding-box residual -- BoundingBoxFactor::operator() (bounding_box_factor.h:68-136) via
// getCornerLocationsVectorRectified (ellipsoid_utils.h:159-273).  ell = (x y z yaw dx dy dz);
// A4 = (Sigma^-1)^(1/2) diag(fx,fx,fy,fy) row-major; brect = rectified (xmin,xmax,ymin,ymax)
// (bounding_box_factor.cpp:26-39).  Invalid case (an inner sqrt argument <= 0): r = invalid_err, J = 0.
// Jo is 4x7 row-major, Jp is 4x6 (either may be null for a residual-only evaluation).
OBVI_HD void bbox_residual_jacobian(const PoseCam& pc, const double* ell, const double* A4, const double* brect,
                                    double invalid_err, double* r, double* Jo, double* Jp) {
  const double hz = sin(0.5 * ell[3]), hw = cos(0.5 * ell[3]);
  const double cpsi = 1.0 - 2.0 * hz * hz, spsi = 2.0 * hz * hw;  // Quaternion(AngleAxis(yaw, z)).toRotationMatrix()
  double A[9];  // Rcw Rz
  for (int i = 0; i < 3; i++) {
    A[3 * i] = pc.Rcw[3 * i] * cpsi + pc.Rcw[3 * i + 1] * spsi;
    A[3 * i + 1] = -pc.Rcw[3 * i] * spsi + pc.Rcw[3 * i + 1] * cpsi;
    A[3 * i + 2] = pc.Rcw[3 * i + 2];
  }
  double tau[3];
  for (int i = 0; i < 3; i++) tau[i] = pc.Rcw[3 * i] * ell[0] + pc.Rcw[3 * i + 1] * ell[1] + pc.Rcw[3 * i + 2] * ell[2] + pc.tcw[i];
  const double d[3] = {0.25 * ell[4] * ell[4] + kDimReg, 0.25 * ell[5] * ell[5] + kDimReg, 0.25 * ell[6] * ell[6] + kDimReg};
#define OBVI_Q(i, j) (A[3 * (i)] * d[0] * A[3 * (j)] + A[3 * (i) + 1] * d[1] * A[3 * (j) + 1] + A[3 * (i) + 2] * d[2] * A[3 * (j) + 2] - tau[i] * tau[j])
  const double q11 = OBVI_Q(0, 0), q13 = OBVI_Q(0, 2), q22 = OBVI_Q(1, 1), q23 = OBVI_Q(1, 2), q33 = OBVI_Q(2, 2);
#undef OBVI_Q
  const double xin = q13 * q13 - q11 * q33, yin = q23 * q23 - q22 * q33;
  if (xin <= 0.0 || yin <= 0.0) {
    for (int i = 0; i < 4; i++) r[i] = invalid_err;
    if (Jo) for (int i = 0; i < 28; i++) Jo[i] = 0.0;
    if (Jp) for (int i = 0; i < 24; i++) Jp[i] = 0.0;
    return;
  }
  const double xs = sqrt(xin), ys = sqrt(yin), iq = 1.0 / q33;
  const double c4[4] = {(q13 + xs) * iq, (q13 - xs) * iq, (q23 + ys) * iq, (q23 - ys) * iq};
  for (int i = 0; i < 4; i++) {
    double s = 0;
    for (int k = 0; k < 4; k++) s += A4[4 * i + k] * (c4[k] - brect[k]);
    r[i] = s;
  }
  if (!Jo && !Jp) return;
  // forward mode over the 13 parameter directions: (dA, dtau, dd) -> dQ -> d corners -> d r
  for (int dir = 0; dir < 13; dir++) {
    double dA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, dtau[3] = {0, 0, 0}, dd[3] = {0, 0, 0};
    if (dir < 3) {  // ellipsoid centre
      for (int i = 0; i < 3; i++) dtau[i] = pc.Rcw[3 * i + dir];
    } else if (dir == 3) {  // yaw: d Rz = [[-s, -c, 0], [c, -s, 0], [0, 0, 0]]
      for (int i = 0; i < 3; i++) {
        dA[3 * i] = -pc.Rcw[3 * i] * spsi + pc.Rcw[3 * i + 1] * cpsi;
        dA[3 * i + 1] = -pc.Rcw[3 * i] * cpsi - pc.Rcw[3 * i + 1] * spsi;
      }
    } else if (dir < 7) {  // dimensions: d ((x/2)^2 + k) = x / 2
      dd[dir - 4] = 0.5 * ell[dir];
    } else if (dir < 10) {  // pose translation
      for (int i = 0; i < 3; i++) dtau[i] = -pc.Rcw[3 * i + (dir - 7)];
    } else {  // pose rotation: d Rcw = M_k, d tcw = m_k
      const double* Mk = pc.M + 9 * (dir - 10);
      for (int i = 0; i < 3; i++) {
        dA[3 * i] = Mk[3 * i] * cpsi + Mk[3 * i + 1] * spsi;
        dA[3 * i + 1] = -Mk[3 * i] * spsi + Mk[3 * i + 1] * cpsi;
        dA[3 * i + 2] = Mk[3 * i + 2];
        dtau[i] = Mk[3 * i] * ell[0] + Mk[3 * i + 1] * ell[1] + Mk[3 * i + 2] * ell[2] + pc.m[3 * (dir - 10) + i];
      }
    }
#define OBVI_DQ(i, j)                                                                                         \
  ((dA[3 * (i)] * A[3 * (j)] + A[3 * (i)] * dA[3 * (j)]) * d[0] + A[3 * (i)] * A[3 * (j)] * dd[0] +              \
   (dA[3 * (i) + 1] * A[3 * (j) + 1] + A[3 * (i) + 1] * dA[3 * (j) + 1]) * d[1] + A[3 * (i) + 1] * A[3 * (j) + 1] * dd[1] + \
   (dA[3 * (i) + 2] * A[3 * (j) + 2] + A[3 * (i) + 2] * dA[3 * (j) + 2]) * d[2] + A[3 * (i) + 2] * A[3 * (j) + 2] * dd[2] - \
   dtau[i] * tau[j] - tau[i] * dtau[j])
    const double d11 = OBVI_DQ(0, 0), d13 = OBVI_DQ(0, 2), d22 = OBVI_DQ(1, 1), d23 = OBVI_DQ(1, 2), d33 = OBVI_DQ(2, 2);
#undef OBVI_DQ
    const double dxs = (2.0 * q13 * d13 - d11 * q33 - q11 * d33) / (2.0 * xs);
    const double dys = (2.0 * q23 * d23 - d22 * q33 - q22 * d33) / (2.0 * ys);
    const double dc[4] = {(d13 + dxs) * iq - c4[0] * d33 * iq, (d13 - dxs) * iq - c4[1] * d33 * iq,
                          (d23 + dys) * iq - c4[2] * d33 * iq, (d23 - dys) * iq - c4[3] * d33 * iq};
    for (int i = 0; i < 4; i++) {
      const double v = A4[4 * i] * dc[0] + A4[4 * i + 1] * dc[1] + A4[4 * i + 2] * dc[2] + A4[4 * i + 3] * dc[3];
      if (dir < 7) { if (Jo) Jo[7 * i + dir] = v; }
      else { if (Jp) Jp[6 * i + (dir - 7)] = v; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Object-only priors: ShapePriorFactor (shape_prior_factor.h:46-61): r = A3 (ell[4:7] - mean);
// IndependentObjectMapFactor (independent_object_map_factor.h:21-33): r = A7 (ell - mean).
OBVI_HD void shape_residual(const double* ell, const double* A3, const double* mean, double* r) {
  for (int i = 0; i < 3; i++)
    r[i] = A3[3 * i] * (ell[4] - mean[0]) + A3[3 * i + 1] * (ell[5] - mean[1]) + A3[3 * i + 2] * (ell[6] - mean[2]);
}
OBVI_HD void ltm_residual(const double* ell, const double* A7, const double* mean, double* r) {
  for (int i = 0; i < 7; i++) {
    double s = 0;
    for (int k = 0; k < 7; k++) s += A7[7 * i + k] * (ell[k] - mean[k]);
    r[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Relative-pose residual -- RelativePoseFactor::operator() (relative_pose_factor.h:32-61) with
// PoseArrayToAffine (vslam_math_util.h:121-141; identity iff |w| < 1e-8) and Eigen's
// AngleAxis(Matrix3) = AngleAxis(Quaternion(Matrix3)).  The rotation part is differentiated with a
// one-direction dual number pushed through the same quaternion / atan2 formulas, so every branch
// (including the constant angle-0 branch) behaves as under Ceres autodiff.
struct Dual1 { double a, d; };
OBVI_HD Dual1 operator+(Dual1 x, Dual1 y) { return {x.a + y.a, x.d + y.d}; }
OBVI_HD Dual1 operator-(Dual1 x, Dual1 y) { return {x.a - y.a, x.d - y.d}; }
OBVI_HD Dual1 operator*(Dual1 x, Dual1 y) { return {x.a * y.a, x.a * y.d + x.d * y.a}; }
OBVI_HD Dual1 operator/(Dual1 x, Dual1 y) { const double q = x.a / y.a; return {q, (x.d - q * y.d) / y.a}; }
OBVI_HD Dual1 dsqrt(Dual1 x) { const double s = sqrt(x.a); return {s, 0.5 * x.d / s}; }
OBVI_HD Dual1 dconst(double c) { return {c, 0.0}; }

// angle * axis of Eigen::AngleAxis(Matrix3 Re), value and directional derivative
OBVI_HD void log_rotation_eigen(const Dual1* Re, Dual1* out) {
  Dual1 qw, qv[3];
  Dual1 tr = Re[0] + Re[4] + Re[8];
  if (tr.a > 0.0) {
    Dual1 t = dsqrt(tr + dconst(1.0));
    qw = dconst(0.5) * t;
    t = dconst(0.5) / t;
    qv[0] = (Re[7] - Re[5]) * t; qv[1] = (Re[2] - Re[6]) * t; qv[2] = (Re[3] - Re[1]) * t;
  } else {
    int i = 0;
    if (Re[4].a > Re[0].a) i = 1;
    if (Re[8].a > Re[4 * i].a) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    Dual1 t = dsqrt(Re[4 * i] - Re[4 * j] - Re[4 * k] + dconst(1.0));
    qv[i] = dconst(0.5) * t;
    t = dconst(0.5) / t;
    qw = (Re[3 * k + j] - Re[3 * j + k]) * t;
    qv[j] = (Re[3 * j + i] + Re[3 * i + j]) * t;
    qv[k] = (Re[3 * k + i] + Re[3 * i + k]) * t;
  }
  Dual1 n = dsqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2]);
  if (n.a != 0.0) {
    Dual1 aw = qw.a < 0.0 ? Dual1{-qw.a, -qw.d} : qw;
    const double den = n.a * n.a + aw.a * aw.a;
    Dual1 ang = {2.0 * atan2(n.a, aw.a), 2.0 * (aw.a * n.d - n.a * aw.d) / den};
    if (qw.a < 0.0) n = Dual1{-n.a, -n.d};
    for (int i = 0; i < 3; i++) out[i] = ang * (qv[i] / n);
  } else {
    for (int i = 0; i < 3; i++) out[i] = dconst(0.0);
  }
}

// r (6), J1 (6x6), J2 (6x6) row-major.  tm = measured translation, Rm_inv = inverse of the measured
// rotation matrix, A6 = (Sigma^-1)^(1/2).  J1/J2 may be null.
OBVI_HD void relpose_residual_jacobian(const double* p1, const double* p2, const double* tm, const double* Rm_inv,
                                       const double* A6, double* r, double* J1, double* J2) {
  double R1[9], dR1[27], R2[9], dR2[27];
  const double th1 = sqrt(p1[3] * p1[3] + p1[4] * p1[4] + p1[5] * p1[5]);
  const double th2 = sqrt(p2[3] * p2[3] + p2[4] * p2[4] + p2[5] * p2[5]);
  rodrigues_with_derivs(p1 + 3, th1 < kSmallAngle, R1, dR1);
  rodrigues_with_derivs(p2 + 3, th2 < kSmallAngle, R2, dR2);
  const double dt[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double B[9];  // R2 Rm_inv
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[3 * i + j] = R2[3 * i] * Rm_inv[j] + R2[3 * i + 1] * Rm_inv[3 + j] + R2[3 * i + 2] * Rm_inv[6 + j];
  double Re[9];  // R1^T B
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Re[3 * i + j] = R1[i] * B[j] + R1[3 + i] * B[3 + j] + R1[6 + i] * B[6 + j];
  // un = [R1^T dt - tm ; Log(Re)],  dun (6 x 12): columns t1(0-2) w1(3-5) t2(6-8) w2(9-11)
  double un[6], dun[72];
  for (int i = 0; i < 72; i++) dun[i] = 0.0;
  for (int i = 0; i < 3; i++) {
    un[i] = R1[i] * dt[0] + R1[3 + i] * dt[1] + R1[6 + i] * dt[2] - tm[i];
    for (int j = 0; j < 3; j++) { dun[12 * i + j] = -R1[3 * j + i]; dun[12 * i + 6 + j] = R1[3 * j + i]; }
    for (int k = 0; k < 3; k++) {
      const double* D = dR1 + 9 * k;
      dun[12 * i + 3 + k] = D[i] * dt[0] + D[3 + i] * dt[1] + D[6 + i] * dt[2];
    }
  }
  for (int dir = -1; dir < 6; dir++) {  // dir -1: value only; 0-2: w1; 3-5: w2
    Dual1 Rd[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double dv = 0.0;
        if (dir >= 0 && dir < 3) {
          const double* D = dR1 + 9 * dir;
          dv = D[i] * B[j] + D[3 + i] * B[3 + j] + D[6 + i] * B[6 + j];
        } else if (dir >= 3) {
          const double* D = dR2 + 9 * (dir - 3);
          for (int a = 0; a < 3; a++) {
            const double dB = D[3 * a] * Rm_inv[j] + D[3 * a + 1] * Rm_inv[3 + j] + D[3 * a + 2] * Rm_inv[6 + j];
            dv += R1[3 * a + i] * dB;
          }
        }
        Rd[3 * i + j] = Dual1{Re[3 * i + j], dv};
      }
    Dual1 lg[3];
    log_rotation_eigen(Rd, lg);
    if (dir < 0) { for (int i = 0; i < 3; i++) un[3 + i] = lg[i].a; if (!J1 && !J2) break; }
    else { const int col = dir < 3 ? 3 + dir : 9 + (dir - 3); for (int i = 0; i < 3; i++) dun[12 * (3 + i) + col] = lg[i].d; }
  }
  for (int i = 0; i < 6; i++) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += A6[6 * i + k] * un[k];
    r[i] = s;
  }
  if (!J1 && !J2) return;
  for (int i = 0; i < 6; i++)
    for (int c = 0; c < 12; c++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += A6[6 * i + k] * dun[12 * k + c];
      if (c < 6) { if (J1) J1[6 * i + c] = s; }
      else { if (J2) J2[6 * i + (c - 6)] = s; }
    }
}

// ---------------------------------------------------------------------------------------------
// Small SPD inverse by Cholesky (n <= 7), row-major full storage.  Returns false if not SPD.  Fully unrolled (the factor
// stays in registers) and with ONE division per pivot -- an fp64 division is ~35 instructions on the GPU and the textbook
// form needs N (N + 1) / 2 + 2 N^2 of them in a serial chain: inv = L^-T L^-1 with L^-1 built from the pivot reciprocals.
template <int N>
OBVI_HD bool spd_inverse(const double* A, double* inv) {
  double L[N * N], rd[N];
#pragma unroll
  for (int j = 0; j < N; j++) {
    double dj = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; k++) dj -= L[j * N + k] * L[j * N + k];
    if (!(dj > 0.0)) return false;
    dj = sqrt(dj);
    L[j * N + j] = dj;
    rd[j] = 1.0 / dj;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      double s = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i * N + k] * L[j * N + k];
      L[i * N + j] = s * rd[j];
    }
  }
  // M = L^-1 (lower triangular), column by column: M[c][c] = 1 / L[c][c], M[i][c] = -(sum_{k=c}^{i-1} L[i][k] M[k][c]) / L[i][i]
  double M[N * N];
#pragma unroll
  for (int c = 0; c < N; c++) {
    M[c * N + c] = rd[c];
#pragma unroll
    for (int i = c + 1; i < N; i++) {
      double s = 0.0;
#pragma unroll
      for (int k = c; k < i; k++) s -= L[i * N + k] * M[k * N + c];
      M[i * N + c] = s * rd[i];
    }
  }
  // inv = M^T M (symmetric): inv[a][b] = sum_{k >= max(a, b)} M[k][a] M[k][b]
#pragma unroll
  for (int a = 0; a < N; a++) {
#pragma unroll
    for (int b = a; b < N; b++) {
      double s = 0.0;
#pragma unroll
      for (int k = b; k < N; k++) s += M[k * N + a] * M[k * N + b];
      inv[a * N + b] = s;
      inv[b * N + a] = s;
    }
  }
  return true;
}

// 3x3 SPD inverse by cofactors: one division instead of the ~24 of the Cholesky route (fp64 division costs ~35
// instructions on the GPU and this sits in the per-point path).  SPD is checked through the leading minors.
OBVI_HD bool spd_inverse3_cofactor(const double* A, double* inv) {
  const double c00 = A[4] * A[8] - A[5] * A[5], c01 = A[2] * A[5] - A[1] * A[8], c02 = A[1] * A[5] - A[2] * A[4];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double m2 = A[0] * A[4] - A[1] * A[1];
  if (!(A[0] > 0.0) || !(m2 > 0.0) || !(det > 0.0)) return false;
  const double id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = inv[3] = c01 * id; inv[2] = inv[6] = c02 * id;
  inv[4] = (A[0] * A[8] - A[2] * A[2]) * id; inv[5] = inv[7] = (A[1] * A[2] - A[0] * A[5]) * id;
  inv[8] = m2 * id;
  return true;
}

}  // namespace obvi
