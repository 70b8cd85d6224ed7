// TEST INFRASTRUCTURE ONLY -- see ba_oracle.hpp.  PARITY UNPINNED (no Ceres binary, no reference
// golden vectors); checked against oracle/py_oracle.py.  (The projection model of the reprojection factor
// alone is pinned by the reference's simulated sequences, tests/test_vslam_dataset.py.)
//
// What is restated, and from where (paths relative to /root/reference):
//   * residual functors, evaluated on forward-mode dual numbers exactly as ceres::AutoDiffCostFunction
//     would evaluate the reference's templated functors:
//       reprojection   include/refactoring/factors/reprojection_cost_functor.h:56-93 +
//                      include/refactoring/types/vslam_math_util.h:347-394 +
//                      src/refactoring/factors/reprojection_cost_functor.cpp:5-17
//       bounding box   include/refactoring/factors/bounding_box_factor.h:68-136 +
//                      include/refactoring/types/ellipsoid_utils.h:159-273 +
//                      src/refactoring/factors/bounding_box_factor.cpp:7-40
//       shape prior    include/refactoring/factors/shape_prior_factor.h:46-61
//       LTM prior      include/refactoring/factors/independent_object_map_factor.h:21-33
//       relative pose  include/refactoring/factors/relative_pose_factor.h:32-61 +
//                      include/refactoring/types/vslam_math_util.h:121-141
//   * solver options as set by include/refactoring/optimization/object_pose_graph_optimizer.h:651-672
//     (SPARSE_SCHUR, LM, Ceres defaults otherwise)
//   * Ceres semantics (EXTERNAL knowledge of upstream Ceres 1.14/2.x -- trust_region_minimizer.cc,
//     levenberg_marquardt_strategy.cc, trust_region_step_evaluator.cc, loss_function.cc, corrector.cc;
//     SURVEY.md Appendix B): Huber loss + corrector, Jacobi scaling from the first Jacobian, LM diagonal
//     clamp, Schur elimination of points (3) and objects (7), sparse Cholesky of the reduced camera
//     system with a minimum-degree ordering (stand-in for CHOLMOD), radius update, non-monotonic step
//     acceptance, the three tolerance tests, minimum-cost iterate returned.
#include "ba_oracle.hpp"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------ dual numbers
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; i++) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; }  // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; i++) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; i++) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; i++) r.v[i] = -x.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; i++) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> inline Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; const double inv = 1.0 / y.a; r.a = x.a * inv; for (int i = 0; i < N; i++) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& x, double s) { Jet<N> r = x; r.a += s; return r; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& x) { return x + s; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, double s) { Jet<N> r = x; r.a -= s; return r; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& x) { return (-x) + s; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, double s) { Jet<N> r; r.a = x.a * s; for (int i = 0; i < N; i++) r.v[i] = x.v[i] * s; return r; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& x) { return x * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& x, double s) { return x * (1.0 / s); }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& x) { return Jet<N>(s) / x; }
template <int N> inline Jet<N> sqrt(const Jet<N>& x) { Jet<N> r; r.a = std::sqrt(x.a); const double d = 0.5 / r.a; for (int i = 0; i < N; i++) r.v[i] = x.v[i] * d; return r; }
template <int N> inline Jet<N> sin(const Jet<N>& x) { Jet<N> r; r.a = std::sin(x.a); const double c = std::cos(x.a); for (int i = 0; i < N; i++) r.v[i] = c * x.v[i]; return r; }
template <int N> inline Jet<N> cos(const Jet<N>& x) { Jet<N> r; r.a = std::cos(x.a); const double s = -std::sin(x.a); for (int i = 0; i < N; i++) r.v[i] = s * x.v[i]; return r; }
template <int N> inline Jet<N> atan2(const Jet<N>& y, const Jet<N>& x) { Jet<N> r; r.a = std::atan2(y.a, x.a); const double d = 1.0 / (x.a * x.a + y.a * y.a); for (int i = 0; i < N; i++) r.v[i] = (x.a * y.v[i] - y.a * x.v[i]) * d; return r; }
template <int N> inline Jet<N> abs(const Jet<N>& x) { return x.a < 0 ? -x : x; }
// ceres::pow(Jet, double p): derivative p * x^(p-1)
template <int N> inline Jet<N> pow2(const Jet<N>& x) { Jet<N> r; r.a = x.a * x.a; const double d = 2.0 * x.a; for (int i = 0; i < N; i++) r.v[i] = d * x.v[i]; return r; }
inline double pow2(double x) { return x * x; }
inline double val(double x) { return x; }
template <int N> inline double val(const Jet<N>& x) { return x.a; }
using std::abs; using std::atan2; using std::cos; using std::sin; using std::sqrt;

// ------------------------------------------------------------------------------------ small algebra
template <class T> inline void mat3_mul(const T* A, const T* B, T* C) {  // C = A B (row-major)
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <class TA, class T> inline void mat3c_mul(const TA* A, const T* B, T* C) {  // constant A
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[3 * i + j] = B[j] * A[3 * i] + B[3 + j] * A[3 * i + 1] + B[6 + j] * A[3 * i + 2];
}
template <class T> inline void mat3_vec(const T* A, const T* x, T* y) {
  for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
template <class T> inline void mat3c_vec(const double* A, const T* x, T* y) {
  for (int i = 0; i < 3; i++) y[i] = x[0] * A[3 * i] + x[1] * A[3 * i + 1] + x[2] * A[3 * i + 2];
}

// Eigen::AngleAxis::toRotationMatrix
template <class T> inline void angle_axis_to_rot(const T& angle, const T* axis, T* R) {
  T s = sin(angle), c = cos(angle);
  T sa[3] = {s * axis[0], s * axis[1], s * axis[2]};
  T ca[3] = {(1.0 - c) * axis[0], (1.0 - c) * axis[1], (1.0 - c) * axis[2]};
  T tmp = ca[0] * axis[1]; R[1] = tmp - sa[2]; R[3] = tmp + sa[2];
  tmp = ca[0] * axis[2]; R[2] = tmp + sa[1]; R[6] = tmp - sa[1];
  tmp = ca[1] * axis[2]; R[5] = tmp - sa[0]; R[7] = tmp + sa[0];
  R[0] = ca[0] * axis[0] + c; R[4] = ca[1] * axis[1] + c; R[8] = ca[2] * axis[2] + c;
}
template <class T> inline void set_identity(T* R) { for (int i = 0; i < 9; i++) R[i] = T(i % 4 == 0 ? 1.0 : 0.0); }

constexpr double kSmallAngle = 1e-8;                         // vslam_math_util.h:17
const double kDimReg = static_cast<double>(1e-3f);           // ellipsoid_utils.h:22 (a float constant)

// inverse robot rotation as built inline by the reprojection / bbox functors (strict '>' test,
// constant identity otherwise): vslam_math_util.h:361-369, ellipsoid_utils.h:176-184
template <class T> inline void functor_inverse_rotation(const T* w, T* R) {
  T ang = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (val(ang) > kSmallAngle) {
    T axis[3] = {w[0] / ang, w[1] / ang, w[2] / ang};
    angle_axis_to_rot(-ang, axis, R);
  } else {
    set_identity(R);
  }
}
// PoseArrayToAffine (vslam_math_util.h:121-141): identity iff |w| < 1e-8
template <class T> inline void pose_array_rotation(const T* w, T* R) {
  T ang = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (val(ang) < kSmallAngle) {
    set_identity(R);
  } else {
    T axis[3] = {w[0] / ang, w[1] / ang, w[2] / ang};
    angle_axis_to_rot(ang, axis, R);
  }
}

// symmetric Jacobi eigen-decomposition -> (cov^-1)^(1/2), principal root (Eigen: cov.inverse().sqrt())
void sqrt_information(const double* cov, int n, double* out) {
  double A[49], V[49];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { A[i * n + j] = 0.5 * (cov[i * n + j] + cov[j * n + i]); V[i * n + j] = (i == j); }
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0; for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      if (A[p * n + q] == 0.0) continue;
      double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * A[p * n + q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < n; k++) { double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; k++) { double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
    }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
    double s = 0; for (int k = 0; k < n; k++) s += V[i * n + k] * V[j * n + k] / std::sqrt(A[k * n + k]);
    out[i * n + j] = s;
  }
}

void inverse3(const double* m, double* o) {
  double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// ------------------------------------------------------------------------------------ functors
struct CamInv { double Rinv[9]; double tinv[3]; double fx, fy, cx, cy; };  // (Translation(t) * R).inverse()

struct ReprojFunctor {
  double rect_x, rect_y, mult_x, mult_y;
  const CamInv* cam;
  template <class T> void operator()(const T* pose, const T* point, T* res) const {
    T Rinv[9]; functor_inverse_rotation(pose + 3, Rinv);
    T Rt[3]; mat3_vec(Rinv, pose, Rt);
    T tr[3] = {-Rt[0], -Rt[1], -Rt[2]};
    T lin[9]; mat3c_mul(cam->Rinv, Rinv, lin);
    T trn[3]; mat3c_vec(cam->Rinv, tr, trn);
    for (int i = 0; i < 3; i++) trn[i] = trn[i] + cam->tinv[i];
    T pc[3]; mat3_vec(lin, point, pc);
    for (int i = 0; i < 3; i++) pc[i] = pc[i] + trn[i];
    res[0] = mult_x * (pc[0] / pc[2] - rect_x);
    res[1] = mult_y * (pc[1] / pc[2] - rect_y);
  }
};

struct BBoxFunctor {
  double A[16];      // (Sigma^-1)^(1/2) diag(fx,fx,fy,fy)
  double brect[4];
  double invalid_err;
  const CamInv* cam;
  template <class T> void operator()(const T* ell, const T* pose, T* res) const {
    T Rinv[9]; functor_inverse_rotation(pose + 3, Rinv);
    T Rt[3]; mat3_vec(Rinv, pose, Rt);
    T tr[3] = {-Rt[0], -Rt[1], -Rt[2]};
    T Rcw[9]; mat3c_mul(cam->Rinv, Rinv, Rcw);
    T tcw[3]; mat3c_vec(cam->Rinv, tr, tcw);
    for (int i = 0; i < 3; i++) tcw[i] = tcw[i] + cam->tinv[i];
    T d[4] = {pow2(ell[4] / 2.0) + kDimReg, pow2(ell[5] / 2.0) + kDimReg, pow2(ell[6] / 2.0) + kDimReg, T(-1.0)};
    // Quaternion(AngleAxis(yaw, UnitZ)).toRotationMatrix()
    T ha = ell[3] * 0.5; T qw = cos(ha), qz = sin(ha);
    T tz = 2.0 * qz; T twz = tz * qw, tzz = tz * qz;
    T Rz[9] = {1.0 - tzz, -twz, T(0.0), twz, 1.0 - tzz, T(0.0), T(0.0), T(0.0), T(1.0)};
    T M[12];  // 3x4 = [Rcw Rz | Rcw c + tcw]
    T L[9]; mat3_mul(Rcw, Rz, L);
    T tc[3]; mat3_vec(Rcw, ell, tc);
    for (int i = 0; i < 3; i++) { M[4 * i] = L[3 * i]; M[4 * i + 1] = L[3 * i + 1]; M[4 * i + 2] = L[3 * i + 2]; M[4 * i + 3] = tc[i] + tcw[i]; }
    auto q = [&](int i, int j) { T s = M[4 * i] * d[0] * M[4 * j]; for (int k = 1; k < 4; k++) s = s + M[4 * i + k] * d[k] * M[4 * j + k]; return s; };
    T q11 = q(0, 0), q13 = q(0, 2), q22 = q(1, 1), q23 = q(1, 2), q33 = q(2, 2);
    T xin = pow2(q13) - q11 * q33, yin = pow2(q23) - q22 * q33;
    if (val(xin) <= 0.0 || val(yin) <= 0.0) { for (int i = 0; i < 4; i++) res[i] = T(invalid_err); return; }
    T xs = sqrt(xin), ys = sqrt(yin);
    T c[4] = {(q13 + xs) / q33, (q13 - xs) / q33, (q23 + ys) / q33, (q23 - ys) / q33};
    T dev[4]; for (int i = 0; i < 4; i++) dev[i] = c[i] - brect[i];
    for (int i = 0; i < 4; i++) { T s = dev[0] * A[4 * i]; for (int k = 1; k < 4; k++) s = s + dev[k] * A[4 * i + k]; res[i] = s; }
  }
};

struct ShapeFunctor {
  double A[9], mean[3];
  template <class T> void operator()(const T* ell, T* res) const {
    T dev[3] = {ell[4] - mean[0], ell[5] - mean[1], ell[6] - mean[2]};
    for (int i = 0; i < 3; i++) res[i] = dev[0] * A[3 * i] + dev[1] * A[3 * i + 1] + dev[2] * A[3 * i + 2];
  }
};
struct LtmFunctor {
  double A[49], mean[7];
  template <class T> void operator()(const T* ell, T* res) const {
    for (int i = 0; i < 7; i++) { T s = (ell[0] - mean[0]) * A[7 * i]; for (int k = 1; k < 7; k++) s = s + (ell[k] - mean[k]) * A[7 * i + k]; res[i] = s; }
  }
};
struct RelPoseFunctor {
  double A[36], tm[3], Rm_inv[9];
  template <class T> void operator()(const T* p1, const T* p2, T* res) const {
    T R1[9], R2[9]; pose_array_rotation(p1 + 3, R1); pose_array_rotation(p2 + 3, R2);
    T R1t[9] = {R1[0], R1[3], R1[6], R1[1], R1[4], R1[7], R1[2], R1[5], R1[8]};
    T dt[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    T t12[3]; mat3_vec(R1t, dt, t12);
    T R12[9]; mat3_mul(R1t, R2, R12);
    T Re[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Re[3 * i + j] = R12[3 * i] * Rm_inv[j] + R12[3 * i + 1] * Rm_inv[3 + j] + R12[3 * i + 2] * Rm_inv[6 + j];
    // Eigen::Quaternion(Matrix3)
    T qw, qv[3];
    T tr = Re[0] + Re[4] + Re[8];
    if (val(tr) > 0.0) {
      T t = sqrt(tr + 1.0); qw = 0.5 * t; t = 0.5 / t;
      qv[0] = (Re[7] - Re[5]) * t; qv[1] = (Re[2] - Re[6]) * t; qv[2] = (Re[3] - Re[1]) * t;
    } else {
      int i = 0; if (val(Re[4]) > val(Re[0])) i = 1; if (val(Re[8]) > val(Re[4 * i])) i = 2;
      int j = (i + 1) % 3, k = (j + 1) % 3;
      T t = sqrt(Re[4 * i] - Re[4 * j] - Re[4 * k] + 1.0);
      qv[i] = 0.5 * t; t = 0.5 / t;
      qw = (Re[3 * k + j] - Re[3 * j + k]) * t;
      qv[j] = (Re[3 * j + i] + Re[3 * i + j]) * t;
      qv[k] = (Re[3 * k + i] + Re[3 * i + k]) * t;
    }
    // Eigen::AngleAxis = Quaternion
    T un[6] = {t12[0] - tm[0], t12[1] - tm[1], t12[2] - tm[2], T(0.0), T(0.0), T(0.0)};
    T n = sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2]);
    if (val(n) != 0.0) {
      T ang = 2.0 * atan2(n, abs(qw));
      if (val(qw) < 0.0) n = -n;
      for (int i = 0; i < 3; i++) un[3 + i] = ang * (qv[i] / n);
    }
    for (int i = 0; i < 6; i++) { T s = un[0] * A[6 * i]; for (int k = 1; k < 6; k++) s = s + un[k] * A[6 * i + k]; res[i] = s; }
  }
};

// autodiff of a 2-block functor: KR residuals, N0 + N1 parameters; Jacobians row-major
template <int KR, int N0, int N1, class F>
inline void autodiff2(const F& f, const double* x0, const double* x1, double* r, double* J0, double* J1) {
  typedef Jet<N0 + N1> JT;
  JT a[N0], b[N1 > 0 ? N1 : 1], out[KR];
  for (int i = 0; i < N0; i++) a[i] = JT(x0[i], i);
  for (int i = 0; i < N1; i++) b[i] = JT(x1[i], N0 + i);
  f(a, b, out);
  for (int k = 0; k < KR; k++) {
    r[k] = out[k].a;
    if (J0) for (int i = 0; i < N0; i++) J0[k * N0 + i] = out[k].v[i];
    if (J1) for (int i = 0; i < N1; i++) J1[k * N1 + i] = out[k].v[N0 + i];
  }
}
template <int KR, int N0, class F>
inline void autodiff1(const F& f, const double* x0, double* r, double* J0) {
  typedef Jet<N0> JT;
  JT a[N0], out[KR];
  for (int i = 0; i < N0; i++) a[i] = JT(x0[i], i);
  f(a, out);
  for (int k = 0; k < KR; k++) { r[k] = out[k].a; if (J0) for (int i = 0; i < N0; i++) J0[k * N0 + i] = out[k].v[i]; }
}

// Ceres HuberLoss::Evaluate + Corrector (rho'' <= 0 => plain sqrt(rho') scaling). Returns 0.5*rho.
inline double huber_correct(double a, int k, double* r, double* J0, int n0, double* J1, int n1, bool apply) {
  double s = 0; for (int i = 0; i < k; i++) s += r[i] * r[i];
  if (!apply) return 0.5 * s;
  const double b = a * a;
  if (s <= b) return 0.5 * s;
  const double rt = std::sqrt(s);
  const double rho1 = std::max(std::numeric_limits<double>::min(), a / rt);
  const double sc = std::sqrt(rho1);
  for (int i = 0; i < k; i++) r[i] *= sc;
  if (J0) for (int i = 0; i < k * n0; i++) J0[i] *= sc;
  if (J1) for (int i = 0; i < k * n1; i++) J1[i] *= sc;
  return 0.5 * (2.0 * a * rt - b);
}

// ------------------------------------------------------------------------------------ problem
struct Problem {
  const oracle_graph* g;
  std::vector<CamInv> cams;
  std::vector<ReprojFunctor> f_rp;
  std::vector<BBoxFunctor> f_bb;
  std::vector<ShapeFunctor> f_sh;
  std::vector<LtmFunctor> f_lt;
  std::vector<RelPoseFunctor> f_rl;
  // residual / Jacobian storage (Ceres layout per block)
  std::vector<double> r_rp, jp_rp, jl_rp, r_bb, jo_bb, jp_bb, r_sh, j_sh, r_lt, j_lt, r_rl, j1_rl, j2_rl;
  // Internal block order: like Ceres' reordering of the program for Schur-type solvers, the residual
  // blocks of one e-block (point / object) are made contiguous, e-blocks in order of their first pose.
  // ord_rp[internal] = user index.
  std::vector<int64_t> ord_rp, ord_bb;
  std::vector<int> point_rank, obj_rank;  // processing rank of each point / object

  void build_order() {
    const oracle_graph& G = *g;
    auto make = [&](int64_t n, const int32_t* e_idx, const int32_t* pose_idx, const int32_t* cam_idx, int ne, std::vector<int>& rank, std::vector<int64_t>& ord) {
      std::vector<int> first(ne, 1 << 30);
      for (int64_t i = 0; i < n; i++) first[e_idx[i]] = std::min(first[e_idx[i]], (int)pose_idx[i]);
      std::vector<int> ids(ne); std::iota(ids.begin(), ids.end(), 0);
      std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return first[a] < first[b]; });
      rank.resize(ne); for (int i = 0; i < ne; i++) rank[ids[i]] = i;
      ord.resize(n); std::iota(ord.begin(), ord.end(), (int64_t)0);
      std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) {
        const int ra = rank[e_idx[a]], rb = rank[e_idx[b]];
        if (ra != rb) return ra < rb;
        if (pose_idx[a] != pose_idx[b]) return pose_idx[a] < pose_idx[b];
        return cam_idx[a] < cam_idx[b]; });
    };
    make(G.n_reproj, G.rp_point, G.rp_pose, G.rp_cam, G.P, point_rank, ord_rp);
    make(G.n_bbox, G.bb_obj, G.bb_pose, G.bb_cam, G.O, obj_rank, ord_bb);
  }

  void build_functors() {
    const oracle_graph& G = *g;
    build_order();
    cams.resize(G.C);
    for (int c = 0; c < G.C; c++) {
      const double* R = G.cam_R + 9 * c; const double* t = G.cam_t + 3 * c;
      CamInv& ci = cams[c];
      // (Translation(t) * R).inverse(): linear = R^-1 (R is a rotation -> transpose), translation = -R^-1 t
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ci.Rinv[3 * i + j] = R[3 * j + i];
      for (int i = 0; i < 3; i++) ci.tinv[i] = -(ci.Rinv[3 * i] * t[0] + ci.Rinv[3 * i + 1] * t[1] + ci.Rinv[3 * i + 2] * t[2]);
      ci.fx = G.cam_intr[4 * c]; ci.fy = G.cam_intr[4 * c + 1]; ci.cx = G.cam_intr[4 * c + 2]; ci.cy = G.cam_intr[4 * c + 3];
    }
    f_rp.resize(G.n_reproj);
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < G.n_reproj; n++) {
      const int64_t u = ord_rp[n];
      const CamInv& ci = cams[G.rp_cam[u]];
      ReprojFunctor& f = f_rp[n];
      f.cam = &ci;
      f.rect_x = (G.rp_px[2 * u] - ci.cx) / ci.fx; f.rect_y = (G.rp_px[2 * u + 1] - ci.cy) / ci.fy;
      f.mult_x = ci.fx / G.rp_sigma[u]; f.mult_y = ci.fy / G.rp_sigma[u];
    }
    f_bb.resize(G.n_bbox);
    for (int64_t n = 0; n < G.n_bbox; n++) {
      const int64_t u = ord_bb[n];
      const CamInv& ci = cams[G.bb_cam[u]];
      BBoxFunctor& f = f_bb[n];
      f.cam = &ci; f.invalid_err = G.bb_invalid;
      double sq[16]; sqrt_information(G.bb_cov + 16 * u, 4, sq);
      const double sc[4] = {ci.fx, ci.fx, ci.fy, ci.fy};
      for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) f.A[4 * i + j] = sq[4 * i + j] * sc[j];
      const double* c4 = G.bb_corners + 4 * u;
      f.brect[0] = (c4[0] - ci.cx) / ci.fx; f.brect[1] = (c4[1] - ci.cx) / ci.fx;
      f.brect[2] = (c4[2] - ci.cy) / ci.fy; f.brect[3] = (c4[3] - ci.cy) / ci.fy;
    }
    f_sh.resize(G.n_shape);
    for (int64_t n = 0; n < G.n_shape; n++) { sqrt_information(G.sh_cov + 9 * n, 3, f_sh[n].A); std::memcpy(f_sh[n].mean, G.sh_mean + 3 * n, 24); }
    f_lt.resize(G.n_ltm);
    for (int64_t n = 0; n < G.n_ltm; n++) { sqrt_information(G.lt_cov + 49 * n, 7, f_lt[n].A); std::memcpy(f_lt[n].mean, G.lt_mean + 7 * n, 56); }
    f_rl.resize(G.n_rel);
    for (int64_t n = 0; n < G.n_rel; n++) {
      sqrt_information(G.rl_cov + 36 * n, 6, f_rl[n].A);
      std::memcpy(f_rl[n].tm, G.rl_t + 3 * n, 24);
      inverse3(G.rl_R + 9 * n, f_rl[n].Rm_inv);
    }
  }
  void alloc(bool jac) {
    const oracle_graph& G = *g;
    r_rp.resize(2 * G.n_reproj); r_bb.resize(4 * G.n_bbox); r_sh.resize(3 * G.n_shape); r_lt.resize(7 * G.n_ltm); r_rl.resize(6 * G.n_rel);
    if (jac) {
      jp_rp.resize(12 * G.n_reproj); jl_rp.resize(6 * G.n_reproj); jo_bb.resize(28 * G.n_bbox); jp_bb.resize(24 * G.n_bbox);
      j_sh.resize(21 * G.n_shape); j_lt.resize(49 * G.n_ltm); j1_rl.resize(36 * G.n_rel); j2_rl.resize(36 * G.n_rel);
    }
  }

  // Evaluate all blocks at (poses, points, objects). Returns total cost (incl. blocks with only constant
  // parameters); `skip` (optional, per type) marks blocks excluded from the cost (fixed-cost blocks).
  double evaluate(const double* poses, const double* points, const double* objs, bool jac, bool apply_loss,
                  double* fixed_cost_out = nullptr) {
    const oracle_graph& G = *g;
    double cost = 0, fixed = 0;
#pragma omp parallel for schedule(static) reduction(+ : cost, fixed)
    for (int64_t n = 0; n < G.n_reproj; n++) {
      const int64_t u = ord_rp[n];
      const double* pp = poses + 6 * G.rp_pose[u]; const double* xp = points + 3 * G.rp_point[u];
      double rloc[2]; double* r = jac ? &r_rp[2 * n] : rloc;
      double c;
      if (jac) {
        double* J0 = &jp_rp[12 * n]; double* J1 = &jl_rp[6 * n];
        autodiff2<2, 6, 3>(f_rp[n], pp, xp, r, J0, J1);
        c = huber_correct(G.rp_huber, 2, r, J0, 6, J1, 3, apply_loss);
      } else {
        f_rp[n](pp, xp, r);
        c = huber_correct(G.rp_huber, 2, r, nullptr, 0, nullptr, 0, apply_loss);
      }
      if (G.const_pose[G.rp_pose[u]] && G.const_point[G.rp_point[u]]) fixed += c; else cost += c;
    }
#pragma omp parallel for schedule(static) reduction(+ : cost, fixed)
    for (int64_t n = 0; n < G.n_bbox; n++) {
      const int64_t u = ord_bb[n];
      const double* e = objs + 7 * G.bb_obj[u]; const double* pp = poses + 6 * G.bb_pose[u];
      double rloc[4]; double* r = jac ? &r_bb[4 * n] : rloc;
      double c;
      if (jac) {
        double* J0 = &jo_bb[28 * n]; double* J1 = &jp_bb[24 * n];
        autodiff2<4, 7, 6>(f_bb[n], e, pp, r, J0, J1);
        c = huber_correct(G.bb_huber, 4, r, J0, 7, J1, 6, apply_loss);
      } else {
        f_bb[n](e, pp, r);
        c = huber_correct(G.bb_huber, 4, r, nullptr, 0, nullptr, 0, apply_loss);
      }
      if (G.const_obj[G.bb_obj[u]] && G.const_pose[G.bb_pose[u]]) fixed += c; else cost += c;
    }
    for (int64_t n = 0; n < G.n_shape; n++) {
      const double* e = objs + 7 * G.sh_obj[n]; double rloc[3]; double* r = jac ? &r_sh[3 * n] : rloc; double c;
      if (jac) { autodiff1<3, 7>(f_sh[n], e, r, &j_sh[21 * n]); c = huber_correct(G.sh_huber, 3, r, &j_sh[21 * n], 7, nullptr, 0, apply_loss); }
      else { f_sh[n](e, r); c = huber_correct(G.sh_huber, 3, r, nullptr, 0, nullptr, 0, apply_loss); }
      if (G.const_obj[G.sh_obj[n]]) fixed += c; else cost += c;
    }
    for (int64_t n = 0; n < G.n_ltm; n++) {
      const double* e = objs + 7 * G.lt_obj[n]; double rloc[7]; double* r = jac ? &r_lt[7 * n] : rloc; double c;
      if (jac) { autodiff1<7, 7>(f_lt[n], e, r, &j_lt[49 * n]); c = huber_correct(G.lt_huber, 7, r, &j_lt[49 * n], 7, nullptr, 0, apply_loss); }
      else { f_lt[n](e, r); c = huber_correct(G.lt_huber, 7, r, nullptr, 0, nullptr, 0, apply_loss); }
      if (G.const_obj[G.lt_obj[n]]) fixed += c; else cost += c;
    }
    for (int64_t n = 0; n < G.n_rel; n++) {
      const double* a = poses + 6 * G.rl_p1[n]; const double* b = poses + 6 * G.rl_p2[n]; double rloc[6]; double* r = jac ? &r_rl[6 * n] : rloc; double c;
      if (jac) { autodiff2<6, 6, 6>(f_rl[n], a, b, r, &j1_rl[36 * n], &j2_rl[36 * n]); c = huber_correct(G.rl_huber, 6, r, &j1_rl[36 * n], 6, &j2_rl[36 * n], 6, apply_loss); }
      else { f_rl[n](a, b, r); c = huber_correct(G.rl_huber, 6, r, nullptr, 0, nullptr, 0, apply_loss); }
      if (G.const_pose[G.rl_p1[n]] && G.const_pose[G.rl_p2[n]]) fixed += c; else cost += c;
    }
    if (fixed_cost_out) *fixed_cost_out = fixed;
    return cost;
  }
};

// ------------------------------------------------------------------------------------ reduced program
// One row record per residual block that touches at least one variable block.
struct Row {
  int k;            // residual size
  int e;            // e-block id (-1: none)
  int f1, f2;       // f-block (variable pose) ids, -1: none
  double *r, *E, *F1, *F2;  // pointers into the Problem's storage
};

struct EBlock { int size; int kind; int idx; int col; };  // kind 0 point, 1 object; col = offset in the reduced vector

struct Reduced {
  int nf = 0, ne = 0, ncols = 0, ecol0 = 0;
  std::vector<int> pose_of_f, f_of_pose;
  std::vector<EBlock> eb;
  std::vector<int> e_of_point, e_of_obj;
  std::vector<Row> rows;
  std::vector<int64_t> erow_ptr; std::vector<int> erow;  // e-block -> rows
  std::vector<int> frows;                                  // rows without an e-block
  std::vector<int> e_order;                                // e-blocks sorted by first f-block (cache locality in S)
  // S structure: upper block CSR over f-blocks
  std::vector<int> s_ptr, s_col;
  // elimination ordering + symbolic factor
  std::vector<int> perm, iperm;            // perm[new] = old
  std::vector<int> l_ptr, l_row;           // block CSC of L (strictly below diagonal), new indices
};

inline int s_find(const Reduced& R, int i, int j) {  // position of block (i,j), i<=j
  const int* b = &R.s_col[R.s_ptr[i]]; const int* e = &R.s_col[R.s_ptr[i + 1]];
  const int* p = std::lower_bound(b, e, j);
  return (int)(p - &R.s_col[0]);
}

void build_reduced(Problem& P, Reduced& R) {
  const oracle_graph& G = *P.g;
  std::vector<char> pose_used(G.K, 0), point_used(G.P, 0), obj_used(G.O, 0);
  for (int64_t n = 0; n < G.n_reproj; n++) { pose_used[G.rp_pose[n]] = 1; point_used[G.rp_point[n]] = 1; }
  for (int64_t n = 0; n < G.n_bbox; n++) { pose_used[G.bb_pose[n]] = 1; obj_used[G.bb_obj[n]] = 1; }
  for (int64_t n = 0; n < G.n_shape; n++) obj_used[G.sh_obj[n]] = 1;
  for (int64_t n = 0; n < G.n_ltm; n++) obj_used[G.lt_obj[n]] = 1;
  for (int64_t n = 0; n < G.n_rel; n++) { pose_used[G.rl_p1[n]] = 1; pose_used[G.rl_p2[n]] = 1; }
  R.f_of_pose.assign(G.K, -1);
  for (int k = 0; k < G.K; k++) if (pose_used[k] && !G.const_pose[k]) { R.f_of_pose[k] = R.nf++; R.pose_of_f.push_back(k); }
  R.e_of_point.assign(G.P, -1); R.e_of_obj.assign(G.O, -1);
  int col = 6 * R.nf; R.ecol0 = col;
  { std::vector<int> by_rank(G.P); for (int p = 0; p < G.P; p++) by_rank[P.point_rank[p]] = p;
    for (int p : by_rank) if (point_used[p] && !G.const_point[p]) { R.e_of_point[p] = (int)R.eb.size(); R.eb.push_back({3, 0, p, col}); col += 3; } }
  { std::vector<int> by_rank(G.O); for (int o = 0; o < G.O; o++) by_rank[P.obj_rank[o]] = o;
    for (int o : by_rank) if (obj_used[o] && !G.const_obj[o]) { R.e_of_obj[o] = (int)R.eb.size(); R.eb.push_back({7, 1, o, col}); col += 7; } }
  R.ne = (int)R.eb.size(); R.ncols = col;
  // rows
  for (int64_t n = 0; n < G.n_reproj; n++) {
    int e = R.e_of_point[G.rp_point[P.ord_rp[n]]], f = R.f_of_pose[G.rp_pose[P.ord_rp[n]]];
    if (e < 0 && f < 0) continue;
    R.rows.push_back({2, e, f, -1, &P.r_rp[2 * n], &P.jl_rp[6 * n], &P.jp_rp[12 * n], nullptr});
  }
  for (int64_t n = 0; n < G.n_bbox; n++) {
    int e = R.e_of_obj[G.bb_obj[P.ord_bb[n]]], f = R.f_of_pose[G.bb_pose[P.ord_bb[n]]];
    if (e < 0 && f < 0) continue;
    R.rows.push_back({4, e, f, -1, &P.r_bb[4 * n], &P.jo_bb[28 * n], &P.jp_bb[24 * n], nullptr});
  }
  for (int64_t n = 0; n < G.n_shape; n++) { int e = R.e_of_obj[G.sh_obj[n]]; if (e >= 0) R.rows.push_back({3, e, -1, -1, &P.r_sh[3 * n], &P.j_sh[21 * n], nullptr, nullptr}); }
  for (int64_t n = 0; n < G.n_ltm; n++) { int e = R.e_of_obj[G.lt_obj[n]]; if (e >= 0) R.rows.push_back({7, e, -1, -1, &P.r_lt[7 * n], &P.j_lt[49 * n], nullptr, nullptr}); }
  for (int64_t n = 0; n < G.n_rel; n++) {
    int f1 = R.f_of_pose[G.rl_p1[n]], f2 = R.f_of_pose[G.rl_p2[n]];
    if (f1 < 0 && f2 < 0) continue;
    R.rows.push_back({6, -1, f1, f2, &P.r_rl[6 * n], nullptr, &P.j1_rl[36 * n], &P.j2_rl[36 * n]});
  }
  // e-block -> rows CSR
  R.erow_ptr.assign(R.ne + 1, 0);
  for (const Row& r : R.rows) if (r.e >= 0) R.erow_ptr[r.e + 1]++;
  for (int e = 0; e < R.ne; e++) R.erow_ptr[e + 1] += R.erow_ptr[e];
  R.erow.resize(R.erow_ptr[R.ne]);
  { std::vector<int64_t> cur(R.erow_ptr.begin(), R.erow_ptr.end() - 1);
    for (size_t i = 0; i < R.rows.size(); i++) { const Row& r = R.rows[i]; if (r.e >= 0) R.erow[cur[r.e]++] = (int)i; else R.frows.push_back((int)i); } }
  { std::vector<int> firstf(R.ne, 1 << 30);
    for (const Row& r : R.rows) if (r.e >= 0 && r.f1 >= 0) firstf[r.e] = std::min(firstf[r.e], r.f1);
    R.e_order.resize(R.ne); std::iota(R.e_order.begin(), R.e_order.end(), 0);
    std::stable_sort(R.e_order.begin(), R.e_order.end(), [&](int a, int b) { return firstf[a] < firstf[b]; }); }
  // S structure
  std::vector<std::vector<int>> cols(R.nf);
  for (int i = 0; i < R.nf; i++) cols[i].push_back(i);
  std::vector<int> fs;
  for (int e = 0; e < R.ne; e++) {
    fs.clear();
    for (int64_t q = R.erow_ptr[e]; q < R.erow_ptr[e + 1]; q++) { int f = R.rows[R.erow[q]].f1; if (f >= 0) fs.push_back(f); }
    std::sort(fs.begin(), fs.end()); fs.erase(std::unique(fs.begin(), fs.end()), fs.end());
    for (size_t a = 0; a < fs.size(); a++) for (size_t b = a + 1; b < fs.size(); b++) cols[fs[a]].push_back(fs[b]);
  }
  for (int i : R.frows) { const Row& r = R.rows[i]; if (r.f1 >= 0 && r.f2 >= 0) cols[std::min(r.f1, r.f2)].push_back(std::max(r.f1, r.f2)); }
  R.s_ptr.assign(R.nf + 1, 0);
  for (int i = 0; i < R.nf; i++) { auto& c = cols[i]; std::sort(c.begin(), c.end()); c.erase(std::unique(c.begin(), c.end()), c.end()); R.s_ptr[i + 1] = R.s_ptr[i] + (int)c.size(); }
  R.s_col.resize(R.s_ptr[R.nf]);
  for (int i = 0; i < R.nf; i++) std::copy(cols[i].begin(), cols[i].end(), R.s_col.begin() + R.s_ptr[i]);

  // minimum-degree ordering on the block graph + symbolic factorisation (bitset elimination graph)
  const int n = R.nf, W = (n + 63) / 64;
  std::vector<uint64_t> adj((size_t)n * W, 0);
  auto setb = [&](int i, int j) { adj[(size_t)i * W + (j >> 6)] |= (1ull << (j & 63)); };
  for (int i = 0; i < n; i++) for (int q = R.s_ptr[i]; q < R.s_ptr[i + 1]; q++) { int j = R.s_col[q]; if (j != i) { setb(i, j); setb(j, i); } }
  std::vector<int> deg(n, 0); std::vector<char> done(n, 0);
  for (int i = 0; i < n; i++) { int d = 0; for (int w = 0; w < W; w++) d += __builtin_popcountll(adj[(size_t)i * W + w]); deg[i] = d; }
  R.perm.resize(n); R.iperm.resize(n);
  std::vector<std::vector<int>> lstruct(n);
  std::vector<int> nb;
  for (int step = 0; step < n; step++) {
    int best = -1;
    for (int i = 0; i < n; i++) if (!done[i] && (best < 0 || deg[i] < deg[best])) best = i;
    done[best] = 1; R.perm[step] = best; R.iperm[best] = step;
    nb.clear();
    uint64_t* ab = &adj[(size_t)best * W];
    for (int w = 0; w < W; w++) { uint64_t m = ab[w]; while (m) { int b = __builtin_ctzll(m); nb.push_back(w * 64 + b); m &= m - 1; } }
    lstruct[step] = nb;  // old indices; converted below
    for (int u : nb) {
      uint64_t* au = &adj[(size_t)u * W];
      for (int w = 0; w < W; w++) au[w] |= ab[w];
      au[u >> 6] &= ~(1ull << (u & 63));
      au[best >> 6] &= ~(1ull << (best & 63));
      int d = 0; for (int w = 0; w < W; w++) d += __builtin_popcountll(au[w]); deg[u] = d;
    }
  }
  R.l_ptr.assign(n + 1, 0);
  for (int s = 0; s < n; s++) { auto& v = lstruct[s]; for (int& u : v) u = R.iperm[u]; std::sort(v.begin(), v.end()); R.l_ptr[s + 1] = R.l_ptr[s] + (int)v.size(); }
  R.l_row.resize(R.l_ptr[n]);
  for (int s = 0; s < n; s++) std::copy(lstruct[s].begin(), lstruct[s].end(), R.l_row.begin() + R.l_ptr[s]);
}

// dense SPD helpers (n <= 7), row-major
template <int N> inline bool chol_inplace(double* A) {
  for (int j = 0; j < N; j++) {
    double d = A[j * N + j]; for (int k = 0; k < j; k++) d -= A[j * N + k] * A[j * N + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d); A[j * N + j] = d;
    for (int i = j + 1; i < N; i++) { double s = A[i * N + j]; for (int k = 0; k < j; k++) s -= A[i * N + k] * A[j * N + k]; A[i * N + j] = s / d; }
  }
  return true;
}
inline bool spd_inverse(const double* A, int n, double* inv) {  // via Cholesky
  double L[49]; std::memcpy(L, A, sizeof(double) * n * n);
  for (int j = 0; j < n; j++) {
    double d = L[j * n + j]; for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d); L[j * n + j] = d;
    for (int i = j + 1; i < n; i++) { double s = L[i * n + j]; for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k]; L[i * n + j] = s / d; }
  }
  for (int c = 0; c < n; c++) {  // solve L L^T x = e_c
    double y[7];
    for (int i = 0; i < n; i++) { double s = (i == c); for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k]; y[i] = s / L[i * n + i]; }
    for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * inv[k * n + c]; inv[i * n + c] = s / L[i * n + i]; }
  }
  return true;
}

// Sparse block Cholesky of the reduced camera system. Sval: upper BSR values (36 per block, row-major
// block (i,j) holds S_ij). Solves S y = b in place of b (length 6*nf). Returns false if not SPD.
bool reduced_solve(const Reduced& R, const std::vector<double>& Sval, std::vector<double>& b,
                   std::vector<double>& Lval, std::vector<double>& Ldiag) {
  const int n = R.nf;
  Lval.assign((size_t)R.l_ptr[n] * 36, 0.0);  // block (row r, col s): L_rs stored row-major 6x6
  Ldiag.assign((size_t)n * 36, 0.0);
  // scatter permuted S into the lower-triangular factor storage
  for (int i = 0; i < n; i++) for (int q = R.s_ptr[i]; q < R.s_ptr[i + 1]; q++) {
    int j = R.s_col[q]; const double* v = &Sval[(size_t)q * 36];
    int pi = R.iperm[i], pj = R.iperm[j];
    if (i == j) { std::memcpy(&Ldiag[(size_t)pi * 36], v, 36 * sizeof(double)); continue; }
    int c = std::min(pi, pj), r = std::max(pi, pj);
    const int* bb = &R.l_row[R.l_ptr[c]]; const int* ee = &R.l_row[R.l_ptr[c + 1]];
    size_t pos = std::lower_bound(bb, ee, r) - &R.l_row[0];
    double* d = &Lval[pos * 36];
    if (pi > pj) { for (int a = 0; a < 6; a++) for (int bq = 0; bq < 6; bq++) d[a * 6 + bq] = v[a * 6 + bq]; }      // L_(pi,pj) = S_ij
    else { for (int a = 0; a < 6; a++) for (int bq = 0; bq < 6; bq++) d[a * 6 + bq] = v[bq * 6 + a]; }             // L_(pj,pi) = S_ij^T
  }
  // right-looking factorisation
  for (int s = 0; s < n; s++) {
    double* D = &Ldiag[(size_t)s * 36];
    if (!chol_inplace<6>(D)) return false;
    const int q0 = R.l_ptr[s], q1 = R.l_ptr[s + 1];
    // L_rs = A_rs D^-T
#pragma omp parallel for schedule(static) if (q1 - q0 > 64)
    for (int q = q0; q < q1; q++) {
      double* Bk = &Lval[(size_t)q * 36];
      for (int a = 0; a < 6; a++) for (int j = 0; j < 6; j++) { double v = Bk[a * 6 + j]; for (int k = 0; k < j; k++) v -= Bk[a * 6 + k] * D[j * 6 + k]; Bk[a * 6 + j] = v / D[j * 6 + j]; }
    }
    // trailing update: A_rc -= L_rs L_cs^T for r >= c in struct(s)
#pragma omp parallel for schedule(dynamic, 4) if (q1 - q0 > 32)
    for (int qc = q0; qc < q1; qc++) {
      const int c = R.l_row[qc]; const double* Lc = &Lval[(size_t)qc * 36];
      double* Dc = &Ldiag[(size_t)c * 36];
      for (int a = 0; a < 6; a++) for (int bq = 0; bq <= a; bq++) { double v = 0; for (int k = 0; k < 6; k++) v += Lc[a * 6 + k] * Lc[bq * 6 + k]; Dc[a * 6 + bq] -= v; }
      int pos = R.l_ptr[c]; const int pend = R.l_ptr[c + 1];
      for (int qr = qc + 1; qr < q1; qr++) {
        const int r = R.l_row[qr];
        while (pos < pend && R.l_row[pos] < r) pos++;
        const double* Lr = &Lval[(size_t)qr * 36]; double* T = &Lval[(size_t)pos * 36];
        for (int a = 0; a < 6; a++) for (int bq = 0; bq < 6; bq++) { double v = 0; for (int k = 0; k < 6; k++) v += Lr[a * 6 + k] * Lc[bq * 6 + k]; T[a * 6 + bq] -= v; }
      }
    }
  }
  // permute rhs, forward / backward substitution
  std::vector<double> y((size_t)6 * n);
  for (int i = 0; i < n; i++) for (int a = 0; a < 6; a++) y[6 * R.iperm[i] + a] = b[6 * i + a];
  for (int s = 0; s < n; s++) {
    const double* D = &Ldiag[(size_t)s * 36]; double* ys = &y[6 * s];
    for (int a = 0; a < 6; a++) { double v = ys[a]; for (int k = 0; k < a; k++) v -= D[a * 6 + k] * ys[k]; ys[a] = v / D[a * 6 + a]; }
    for (int q = R.l_ptr[s]; q < R.l_ptr[s + 1]; q++) { const double* Bk = &Lval[(size_t)q * 36]; double* yr = &y[6 * R.l_row[q]]; for (int a = 0; a < 6; a++) { double v = 0; for (int k = 0; k < 6; k++) v += Bk[a * 6 + k] * ys[k]; yr[a] -= v; } }
  }
  for (int s = n - 1; s >= 0; s--) {
    const double* D = &Ldiag[(size_t)s * 36]; double* ys = &y[6 * s];
    for (int q = R.l_ptr[s]; q < R.l_ptr[s + 1]; q++) { const double* Bk = &Lval[(size_t)q * 36]; const double* yr = &y[6 * R.l_row[q]]; for (int k = 0; k < 6; k++) { double v = 0; for (int a = 0; a < 6; a++) v += Bk[a * 6 + k] * yr[a]; ys[k] -= v; } }
    for (int a = 5; a >= 0; a--) { double v = ys[a]; for (int k = a + 1; k < 6; k++) v -= D[k * 6 + a] * ys[k]; ys[a] = v / D[a * 6 + a]; }
  }
  for (int i = 0; i < n; i++) for (int a = 0; a < 6; a++) b[6 * i + a] = y[6 * R.iperm[i] + a];
  return true;
}

inline int row_e_cols(const Row& r) { return r.e >= 0 ? (r.k == 2 ? 3 : 7) : 0; }

struct Workspace {
  std::vector<double> Sval, rhs, Lval, Ldiag;
  std::vector<std::vector<double>> S_tl, rhs_tl;
  std::vector<double> einv, eg;  // per e-block: inverse (49 max) and E^T r (7 max)
};

// Solve (J^T J + D^2) y = J^T r by Schur elimination of the e-blocks.  J is the (already Jacobi-scaled)
// Jacobian held in the rows; D over all reduced columns.  Returns false on a failed factorisation.
bool schur_solve(const Reduced& R, const std::vector<double>& D, std::vector<double>& y, Workspace& W, int nthreads) {
  const int nf = R.nf, nblk = R.s_ptr[nf];
  const double t_in = now_s();
  W.S_tl.resize(nthreads); W.rhs_tl.resize(nthreads);
  W.einv.resize((size_t)R.ne * 49); W.eg.resize((size_t)R.ne * 7);
  bool ok = true;
#pragma omp parallel num_threads(nthreads)
  {
    const int tid = omp_get_thread_num();
    std::vector<double>& S = W.S_tl[tid]; std::vector<double>& rhs = W.rhs_tl[tid];
    S.assign((size_t)nblk * 36, 0.0); rhs.assign((size_t)6 * nf, 0.0);
    std::vector<int> fs; std::vector<double> Wf;  // merged per-f-block W_i = sum F^T E (6 x ne)
#pragma omp for schedule(static)
    for (int eo = 0; eo < R.ne; eo++) {
      const int e = R.e_order[eo];
      const int ne = R.eb[e].size; const int col = R.eb[e].col;
      double ete[49], g[7];
      for (int i = 0; i < ne * ne; i++) ete[i] = 0; for (int i = 0; i < ne; i++) { g[i] = 0; ete[i * ne + i] = D[col + i] * D[col + i]; }
      fs.clear();
      for (int64_t q = R.erow_ptr[e]; q < R.erow_ptr[e + 1]; q++) {
        const Row& r = R.rows[R.erow[q]];
        for (int k = 0; k < r.k; k++) for (int a = 0; a < ne; a++) { const double ea = r.E[k * ne + a]; g[a] += ea * r.r[k]; for (int b = 0; b < ne; b++) ete[a * ne + b] += ea * r.E[k * ne + b]; }
        if (r.f1 >= 0) fs.push_back(r.f1);
      }
      double* inv = &W.einv[(size_t)e * 49];
      if (!spd_inverse(ete, ne, inv)) { ok = false; continue; }
      for (int i = 0; i < ne; i++) W.eg[(size_t)e * 7 + i] = g[i];
      std::sort(fs.begin(), fs.end()); fs.erase(std::unique(fs.begin(), fs.end()), fs.end());
      const int nfs = (int)fs.size();
      Wf.assign((size_t)nfs * 6 * ne, 0.0);
      for (int64_t q = R.erow_ptr[e]; q < R.erow_ptr[e + 1]; q++) {
        const Row& r = R.rows[R.erow[q]];
        if (r.f1 < 0) continue;
        const int slot = (int)(std::lower_bound(fs.begin(), fs.end(), r.f1) - fs.begin());
        double* Wi = &Wf[(size_t)slot * 6 * ne];
        double* Sd = &S[(size_t)s_find(R, r.f1, r.f1) * 36]; double* bi = &rhs[6 * r.f1];
        for (int k = 0; k < r.k; k++) for (int a = 0; a < 6; a++) {
          const double fa = r.F1[k * 6 + a];
          bi[a] += fa * r.r[k];
          for (int b = 0; b < 6; b++) Sd[a * 6 + b] += fa * r.F1[k * 6 + b];
          for (int b = 0; b < ne; b++) Wi[a * ne + b] += fa * r.E[k * ne + b];
        }
      }
      // S_ij -= W_i inv W_j^T, rhs_i -= W_i inv g
      double Z[42], ig[7];
      for (int a = 0; a < ne; a++) { double s = 0; for (int b = 0; b < ne; b++) s += inv[a * ne + b] * g[b]; ig[a] = s; }
      for (int i = 0; i < nfs; i++) {
        const double* Wi = &Wf[(size_t)i * 6 * ne];
        for (int a = 0; a < 6; a++) for (int b = 0; b < ne; b++) { double s = 0; for (int c = 0; c < ne; c++) s += Wi[a * ne + c] * inv[c * ne + b]; Z[a * ne + b] = s; }
        double* bi = &rhs[6 * fs[i]];
        for (int a = 0; a < 6; a++) { double s = 0; for (int c = 0; c < ne; c++) s += Wi[a * ne + c] * ig[c]; bi[a] -= s; }
        int pos = R.s_ptr[fs[i]];
        for (int j = i; j < nfs; j++) {
          while (R.s_col[pos] < fs[j]) pos++;
          const double* Wj = &Wf[(size_t)j * 6 * ne]; double* Sb = &S[(size_t)pos * 36];
          for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) { double s = 0; for (int c = 0; c < ne; c++) s += Z[a * ne + c] * Wj[b * ne + c]; Sb[a * 6 + b] -= s; }
        }
      }
    }
  }
  const bool dbg = getenv("ORACLE_DEBUG") != nullptr; double td = now_s();
  if (dbg) fprintf(stderr, "[oracle] eliminate %.3f s\n", td - t_in);
  if (!ok) return false;
  // rows without an e-block (relative pose; or blocks whose e-parameter is constant)
  W.Sval.assign((size_t)nblk * 36, 0.0); W.rhs.assign((size_t)6 * nf, 0.0);
  for (int ri : R.frows) {
    const Row& r = R.rows[ri];
    const int fi[2] = {r.f1, r.f2}; const double* Fm[2] = {r.F1, r.F2};
    for (int u = 0; u < 2; u++) {
      if (fi[u] < 0) continue;
      for (int k = 0; k < r.k; k++) for (int a = 0; a < 6; a++) W.rhs[6 * fi[u] + a] += Fm[u][k * 6 + a] * r.r[k];
      for (int v = 0; v < 2; v++) {
        if (fi[v] < 0 || fi[u] > fi[v]) continue;
        if (u != v && fi[u] == fi[v]) continue;
        double* Sb = &W.Sval[(size_t)s_find(R, fi[u], fi[v]) * 36];
        for (int k = 0; k < r.k; k++) for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) Sb[a * 6 + b] += Fm[u][k * 6 + a] * Fm[v][k * 6 + b];
      }
    }
  }
  for (int t = 0; t < nthreads; t++) {
    const std::vector<double>& S = W.S_tl[t]; const std::vector<double>& rh = W.rhs_tl[t];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)nblk * 36; i++) W.Sval[i] += S[i];
    for (int i = 0; i < 6 * nf; i++) W.rhs[i] += rh[i];
  }
  for (int i = 0; i < nf; i++) { double* Sd = &W.Sval[(size_t)R.s_ptr[i] * 36]; for (int a = 0; a < 6; a++) Sd[a * 6 + a] += D[6 * i + a] * D[6 * i + a]; }
  if (dbg) { fprintf(stderr, "[oracle] reduce %.3f s\n", now_s() - td); td = now_s(); }
  if (nf > 0 && !reduced_solve(R, W.Sval, W.rhs, W.Lval, W.Ldiag)) return false;
  if (dbg) { fprintf(stderr, "[oracle] cholesky %.3f s (nnz L blocks %d, S blocks %d)\n", now_s() - td, R.l_ptr[nf], nblk); td = now_s(); }
  y.assign(R.ncols, 0.0);
  for (int i = 0; i < 6 * nf; i++) y[i] = W.rhs[i];
  // back-substitution: y_e = inv (g - sum_i W_i^T y_i) = inv (g - sum_rows E^T (F y_f))
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int e = 0; e < R.ne; e++) {
    const int ne = R.eb[e].size; double t[7];
    for (int a = 0; a < ne; a++) t[a] = W.eg[(size_t)e * 7 + a];
    for (int64_t q = R.erow_ptr[e]; q < R.erow_ptr[e + 1]; q++) {
      const Row& r = R.rows[R.erow[q]];
      if (r.f1 < 0) continue;
      for (int k = 0; k < r.k; k++) { double fy = 0; for (int a = 0; a < 6; a++) fy += r.F1[k * 6 + a] * y[6 * r.f1 + a]; for (int a = 0; a < ne; a++) t[a] -= r.E[k * ne + a] * fy; }
    }
    const double* inv = &W.einv[(size_t)e * 49];
    for (int a = 0; a < ne; a++) { double s = 0; for (int b = 0; b < ne; b++) s += inv[a * ne + b] * t[b]; y[R.eb[e].col + a] = s; }
  }
  return true;
}

}  // namespace

extern "C" void oracle_sqrt_information(const double* cov, int n, double* out) { sqrt_information(cov, n, out); }

extern "C" int oracle_evaluate(const oracle_graph* g, int apply_loss, double* cost, double* r_reproj, double* jp_reproj,
                               double* jl_reproj, double* r_bbox, double* jo_bbox, double* jp_bbox, double* r_shape,
                               double* j_shape, double* r_ltm, double* j_ltm, double* r_rel, double* j1_rel, double* j2_rel) {
  Problem P; P.g = g; P.build_functors(); P.alloc(true);
  double fixed = 0;
  double c = P.evaluate(g->poses, g->points, g->objects, true, apply_loss != 0, &fixed);
  if (cost) *cost = c + fixed;
  auto cp = [](double* dst, const std::vector<double>& src) { if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(double)); };
  auto sc = [](double* dst, const std::vector<double>& src, const std::vector<int64_t>& ord, int w) {
    if (!dst) return; for (size_t n = 0; n < ord.size(); n++) std::memcpy(dst + (size_t)ord[n] * w, &src[n * w], sizeof(double) * w); };
  sc(r_reproj, P.r_rp, P.ord_rp, 2); sc(jp_reproj, P.jp_rp, P.ord_rp, 12); sc(jl_reproj, P.jl_rp, P.ord_rp, 6);
  sc(r_bbox, P.r_bb, P.ord_bb, 4); sc(jo_bbox, P.jo_bb, P.ord_bb, 28); sc(jp_bbox, P.jp_bb, P.ord_bb, 24);
  cp(r_shape, P.r_sh); cp(j_shape, P.j_sh); cp(r_ltm, P.r_lt); cp(j_ltm, P.j_lt); cp(r_rel, P.r_rl); cp(j1_rel, P.j1_rl); cp(j2_rel, P.j2_rl);
  return 0;
}

extern "C" int oracle_solve(oracle_graph* g, const oracle_options* opt, oracle_summary* out, double* log, int32_t log_rows) {
  const double t_start = now_s();
  int nthreads = opt->num_threads > 0 ? opt->num_threads : omp_get_max_threads();
  omp_set_num_threads(nthreads);
  Problem P; P.g = g; P.build_functors(); P.alloc(true);
  Reduced R; build_reduced(P, R);
  const oracle_graph& G = *g;
  std::memset(out, 0, sizeof(*out));
  out->num_parameters_reduced = R.ncols; out->num_threads_used = nthreads;
  // working copies of the state (x) and the candidate
  std::vector<double> poses(G.poses, G.poses + 6 * G.K), points(G.points, G.points + 3 * G.P), objs(G.objects, G.objects + 7 * G.O);
  std::vector<double> c_poses = poses, c_points = points, c_objs = objs;
  const int n = R.ncols;
  std::vector<double> scale(n, 1.0), diag(n, 0.0), D(n, 0.0), grad(n, 0.0), y, delta(n, 0.0);
  Workspace W;
  int log_n = 0;
  auto push_log = [&](int it, double cost, double cc, double sn, int ok, double radius, double gmax) {
    if (log && log_n < log_rows) { double* L = log + (size_t)ORACLE_LOG_COLS * log_n; L[0] = it; L[1] = cost; L[2] = cc; L[3] = sn; L[4] = ok; L[5] = radius; L[6] = gmax; }
    log_n++;
  };
  std::vector<std::vector<double>> sq_tl(nthreads), gr_tl(nthreads);
  auto col_sqnorm_and_grad = [&](std::vector<double>& sq, std::vector<double>* gr) {
#pragma omp parallel num_threads(nthreads)
    {
      const int tid = omp_get_thread_num();
      std::vector<double>& sqt = sq_tl[tid]; std::vector<double>& grt = gr_tl[tid];
      sqt.assign(n, 0.0); if (gr) grt.assign(n, 0.0);
#pragma omp for schedule(static)
      for (int64_t i = 0; i < (int64_t)R.rows.size(); i++) {
        const Row& r = R.rows[i];
        const int ne = row_e_cols(r);
        if (r.e >= 0) { const int c0 = R.eb[r.e].col; for (int k = 0; k < r.k; k++) for (int a = 0; a < ne; a++) { const double v = r.E[k * ne + a]; sqt[c0 + a] += v * v; if (gr) grt[c0 + a] += v * r.r[k]; } }
        if (r.f1 >= 0) for (int k = 0; k < r.k; k++) for (int a = 0; a < 6; a++) { const double v = r.F1[k * 6 + a]; sqt[6 * r.f1 + a] += v * v; if (gr) grt[6 * r.f1 + a] += v * r.r[k]; }
        if (r.f2 >= 0) for (int k = 0; k < r.k; k++) for (int a = 0; a < 6; a++) { const double v = r.F2[k * 6 + a]; sqt[6 * r.f2 + a] += v * v; if (gr) grt[6 * r.f2 + a] += v * r.r[k]; }
      }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
      double a = 0, b = 0;
      for (int t = 0; t < nthreads; t++) { a += sq_tl[t][i]; if (gr) b += gr_tl[t][i]; }
      sq[i] = a; if (gr) (*gr)[i] = b;
    }
  };
  auto scale_jacobian = [&]() {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)R.rows.size(); i++) {
      const Row& r = R.rows[i]; const int ne = row_e_cols(r);
      if (r.e >= 0) { const int c0 = R.eb[r.e].col; for (int k = 0; k < r.k; k++) for (int a = 0; a < ne; a++) r.E[k * ne + a] *= scale[c0 + a]; }
      if (r.f1 >= 0) for (int k = 0; k < r.k; k++) for (int a = 0; a < 6; a++) r.F1[k * 6 + a] *= scale[6 * r.f1 + a];
      if (r.f2 >= 0) for (int k = 0; k < r.k; k++) for (int a = 0; a < 6; a++) r.F2[k * 6 + a] *= scale[6 * r.f2 + a];
    }
  };
  auto x_norm = [&](const std::vector<double>& ps, const std::vector<double>& pt, const std::vector<double>& ob) {
    double s = 0;
    for (int f = 0; f < R.nf; f++) for (int a = 0; a < 6; a++) { double v = ps[6 * R.pose_of_f[f] + a]; s += v * v; }
    for (const EBlock& e : R.eb) { const double* b = e.kind == 0 ? &pt[3 * e.idx] : &ob[7 * e.idx]; for (int a = 0; a < e.size; a++) s += b[a] * b[a]; }
    return std::sqrt(s);
  };
  double t_jac = 0, t_res = 0, t_lin = 0;
  double fixed_cost = 0;
  double t0 = now_s();
  double x_cost = P.evaluate(poses.data(), points.data(), objs.data(), true, true, &fixed_cost);
  col_sqnorm_and_grad(diag, &grad);
  for (int i = 0; i < n; i++) scale[i] = 1.0 / (1.0 + std::sqrt(diag[i]));
  scale_jacobian();
  t_jac += now_s() - t0;
  const double t_iter_start = now_s();
  double gmax = 0; for (int i = 0; i < n; i++) gmax = std::max(gmax, std::fabs(grad[i]));
  out->initial_cost = x_cost + fixed_cost; out->fixed_cost = fixed_cost;
  double radius = opt->initial_trust_region_radius, decrease = 2.0;
  push_log(0, x_cost + fixed_cost, 0, 0, 0, radius, gmax);
  double minimum_cost = x_cost;
  int termination = 1;
  // step evaluator
  const int max_nonmono = opt->use_nonmonotonic_steps ? 5 : 0;
  double ev_min = x_cost, ev_cur = x_cost, ev_ref = x_cost, ev_cand = x_cost, acc_ref = 0, acc_cand = 0; int n_nonmono = 0;
  bool reuse_diag = false; int n_invalid = 0; int iter = 0; int lm_steps = 0;
  double xn = x_norm(poses, points, objs);
  if (n == 0 || gmax <= opt->gradient_tolerance) termination = 0;
  else while (true) {
    if (iter >= opt->max_num_iterations) { termination = 1; break; }
    iter++; lm_steps++;
    t0 = now_s();
    if (!reuse_diag) { col_sqnorm_and_grad(diag, nullptr); for (int i = 0; i < n; i++) diag[i] = std::min(std::max(diag[i], 1e-6), 1e32); }
    for (int i = 0; i < n; i++) D[i] = std::sqrt(diag[i] / radius);
    bool ok = schur_solve(R, D, y, W, nthreads);
    if (ok) for (int i = 0; i < n; i++) if (!std::isfinite(y[i])) { ok = false; break; }
    reuse_diag = true;
    double model_change = 0;
    if (ok) {
      for (int i = 0; i < n; i++) y[i] = -y[i];
      double mc = 0;
#pragma omp parallel for schedule(static) reduction(+ : mc)
      for (int64_t i = 0; i < (int64_t)R.rows.size(); i++) {
        const Row& r = R.rows[i]; const int ne = row_e_cols(r);
        for (int k = 0; k < r.k; k++) {
          double m = 0;
          if (r.e >= 0) { const int c0 = R.eb[r.e].col; for (int a = 0; a < ne; a++) m += r.E[k * ne + a] * y[c0 + a]; }
          if (r.f1 >= 0) for (int a = 0; a < 6; a++) m += r.F1[k * 6 + a] * y[6 * r.f1 + a];
          if (r.f2 >= 0) for (int a = 0; a < 6; a++) m += r.F2[k * 6 + a] * y[6 * r.f2 + a];
          mc += m * (r.r[k] + 0.5 * m);
        }
      }
      model_change = -mc;
      ok = model_change > 0.0;
    }
    t_lin += now_s() - t0;
    if (!ok) {
      if (++n_invalid >= 5) { termination = 2; break; }
      radius /= decrease; decrease *= 2.0; reuse_diag = true;  // LevenbergMarquardtStrategy::StepIsInvalid () = StepRejected (0)
      push_log(iter, x_cost + fixed_cost, 0, 0, 0, radius, gmax);
      continue;
    }
    n_invalid = 0;
    for (int i = 0; i < n; i++) delta[i] = y[i] * scale[i];
    c_poses = poses; c_points = points; c_objs = objs;
    for (int f = 0; f < R.nf; f++) for (int a = 0; a < 6; a++) c_poses[6 * R.pose_of_f[f] + a] += delta[6 * f + a];
    for (const EBlock& e : R.eb) { double* b = e.kind == 0 ? &c_points[3 * e.idx] : &c_objs[7 * e.idx]; for (int a = 0; a < e.size; a++) b[a] += delta[e.col + a]; }
    t0 = now_s();
    // residual-only evaluation (does not touch the stored residuals / Jacobians of x)
    double cand_cost = P.evaluate(c_poses.data(), c_points.data(), c_objs.data(), false, true, nullptr);
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
    t_res += now_s() - t0;
    double step_norm = 0; for (int i = 0; i < n; i++) step_norm += delta[i] * delta[i]; step_norm = std::sqrt(step_norm);
    if (step_norm <= opt->parameter_tolerance * (xn + opt->parameter_tolerance)) { termination = 0; break; }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= opt->function_tolerance * x_cost) { termination = 0; break; }
    double rho;
    if (cand_cost >= std::numeric_limits<double>::max()) rho = std::numeric_limits<double>::lowest();
    else rho = std::max((ev_cur - cand_cost) / model_change, (ev_ref - cand_cost) / (acc_ref + model_change));
    if (rho > 1e-3) {
      poses.swap(c_poses); points.swap(c_points); objs.swap(c_objs);
      xn = x_norm(poses, points, objs);
      t0 = now_s();
      x_cost = P.evaluate(poses.data(), points.data(), objs.data(), true, true, nullptr);
      col_sqnorm_and_grad(diag, &grad);  // diag is recomputed on the scaled Jacobian below
      gmax = 0; for (int i = 0; i < n; i++) gmax = std::max(gmax, std::fabs(grad[i]));
      scale_jacobian();
      t_jac += now_s() - t0;
      radius = std::min(opt->max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      decrease = 2.0; reuse_diag = false;
      ev_cur = cand_cost; acc_cand += model_change; acc_ref += model_change;
      if (ev_cur < ev_min) { ev_min = ev_cur; n_nonmono = 0; ev_cand = ev_cur; acc_cand = 0; }
      else { n_nonmono++; if (ev_cur > ev_cand) { ev_cand = ev_cur; acc_cand = 0; } }
      if (n_nonmono == max_nonmono) { ev_ref = ev_cand; acc_ref = acc_cand; }
      if (x_cost < minimum_cost) {
        minimum_cost = x_cost;
        std::memcpy(G.poses, poses.data(), sizeof(double) * 6 * G.K); std::memcpy(G.points, points.data(), sizeof(double) * 3 * G.P);
        std::memcpy(G.objects, objs.data(), sizeof(double) * 7 * G.O);
      }
      push_log(iter, x_cost + fixed_cost, cost_change, step_norm, 1, radius, gmax);
      if (gmax <= opt->gradient_tolerance) { termination = 0; break; }
    } else {
      radius /= decrease; decrease *= 2.0; reuse_diag = true;
      push_log(iter, cand_cost + fixed_cost, cost_change, step_norm, 0, radius, gmax);
    }
    if (radius <= 1e-32) { termination = 0; break; }
  }
  out->termination = termination; out->num_iterations = log_n; out->lm_steps = lm_steps;
  out->final_cost = minimum_cost + fixed_cost;
  out->total_time = now_s() - t_start; out->linear_solver_time = t_lin; out->jacobian_time = t_jac; out->residual_time = t_res;
  (void)t_iter_start;
  return 0;
}
