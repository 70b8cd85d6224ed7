// Helpers shared by the replacement factor headers: conversion of the reference's Eigen-typed factory arguments
// (include/refactoring/types/vslam_basic_types_refactor.h:18-80, vslam_obj_opt_types_refactor.h:15-39) into the plain
// arrays the obvi_ba C ABI takes.  Only element access (operator(), .x(), .toRotationMatrix()) is used, so the headers
// work with Eigen proper and with any stand-in that offers the same accessors.
#ifndef OBVI_FACTOR_COMMON_H_
#define OBVI_FACTOR_COMMON_H_

#include "obvi_ba.h"

namespace obvi_shim {

struct CameraData { double intr[4]; double R[9]; double t[3]; };

template <typename IntrinsicsMat, typename Extrinsics>
inline CameraData makeCamera(const IntrinsicsMat& K, const Extrinsics& extrinsics) {
  CameraData c;
  c.intr[0] = K(0, 0); c.intr[1] = K(1, 1); c.intr[2] = K(0, 2); c.intr[3] = K(1, 2);
  const auto R = extrinsics.orientation_.toRotationMatrix();
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) c.R[3 * i + j] = R(i, j);
    c.t[i] = extrinsics.transl_(i);
  }
  return c;
}

template <int N, typename Mat>
inline void copySquare(const Mat& m, double* out) {
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) out[N * i + j] = m(i, j);
}

inline int registerCamera(obvi_problem* p, const CameraData& c, int* cam_id) { return obvi_camera_add(p, c.intr, c.R, c.t, cam_id); }

}  // namespace obvi_shim
#endif
