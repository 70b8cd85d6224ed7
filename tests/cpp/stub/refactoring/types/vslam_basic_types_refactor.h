// TEST STUB standing in for the reference's include/refactoring/types/vslam_basic_types_refactor.h (which needs Eigen,
// absent from this image).  It offers the same type NAMES with just the accessors the replacement factor headers use.
#pragma once
#include <cstdint>
#include <memory>

namespace mini {
template <typename T, int R, int C>
struct Matrix {
  T v[R * C] = {};
  T& operator()(int i, int j) { return v[i * C + j]; }
  const T& operator()(int i, int j) const { return v[i * C + j]; }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  static Matrix Zero() { return Matrix(); }
};
template <typename T>
struct AngleAxis {  // stores the rotation matrix directly
  Matrix<T, 3, 3> R;
  Matrix<T, 3, 3> toRotationMatrix() const { return R; }
};
}  // namespace mini

namespace vslam_types_refactor {
typedef uint64_t CameraId;
typedef uint64_t FrameId;
typedef uint64_t FeatureId;
template <typename NumType> using CameraIntrinsicsMat = mini::Matrix<NumType, 3, 3>;
template <typename NumType> using PixelCoord = mini::Matrix<NumType, 2, 1>;
template <typename NumType> using Position3d = mini::Matrix<NumType, 3, 1>;
template <typename NumType> using Orientation3D = mini::AngleAxis<NumType>;
template <typename NumType, int MatDim> using Covariance = mini::Matrix<NumType, MatDim, MatDim>;
template <typename NumType> using RawPose3d = mini::Matrix<NumType, 6, 1>;
template <typename NumType>
struct Pose3D {
  Position3d<NumType> transl_;
  Orientation3D<NumType> orientation_;
};
template <typename NumType>
struct Pose3DYawOnly {
  Position3d<NumType> transl_;
  NumType yaw_ = NumType(0);
};
template <typename NumType> using CameraExtrinsics = Pose3D<NumType>;
}  // namespace vslam_types_refactor
