// Host-side constant preparation for the factor records (what the reference's functor constructors do).
#pragma once
#include <cmath>

namespace obvi {

// (cov^-1)^(1/2), principal square root -- Eigen's `cov.inverse().sqrt()` in
// src/refactoring/factors/{bounding_box_factor.cpp:31-33, relative_pose_factor.cpp:13,
// shape_prior_factor.cpp:11, independent_object_map_factor.cpp:7-11}.  For SPD input this is
// V diag(lambda^-1/2) V^T; computed with a cyclic Jacobi eigen-solver (n <= 7).  Returns false when
// the result has a NaN (the reference exits on that, relative_pose_factor.cpp:14-18).
inline bool sqrt_information(const double* cov, int n, double* out) {
  double A[49], V[49];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      A[i * n + j] = 0.5 * (cov[i * n + j] + cov[j * n + i]);
      V[i * n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0.0;
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off == 0.0) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  bool ok = true;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (int k = 0; k < n; k++) s += V[i * n + k] * V[j * n + k] / std::sqrt(A[k * n + k]);
      out[i * n + j] = s;
      if (!(s == s)) ok = false;
    }
  return ok;
}

// General 3x3 inverse (relative_pose_factor.h:50-51 inverts the measured rotation with .inverse()).
inline void inverse3(const double* m, double* o) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double id = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// Inverse extrinsics: (Translation(t_e) * R_e).inverse() = (R_e^T, -R_e^T t_e)
// (reprojection_cost_functor.cpp:9-11, bounding_box_factor.cpp:19-21).
inline void invert_extrinsics(const double* R, const double* t, double* Rinv, double* tinv) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rinv[3 * i + j] = R[3 * j + i];
  for (int i = 0; i < 3; i++) tinv[i] = -(Rinv[3 * i] * t[0] + Rinv[3 * i + 1] * t[1] + Rinv[3 * i + 2] * t[2]);
}

}  // namespace obvi
