// TEST STUB standing in for include/refactoring/types/vslam_obj_opt_types_refactor.h (CONSTRAIN_ELLIPSOID_ORIENTATION on).
#pragma once
#include <refactoring/types/vslam_basic_types_refactor.h>

namespace vslam_types_refactor {
const static int kEllipsoidPoseParameterizationSize = 4;
const static int kEllipsoidParamterizationSize = kEllipsoidPoseParameterizationSize + 3;
typedef uint64_t ObjectId;
template <typename NumType> using BbCorners = mini::Matrix<NumType, 4, 1>;
template <typename NumType> using ObjectDim = mini::Matrix<NumType, 3, 1>;
template <typename NumType> using RawEllipsoid = mini::Matrix<NumType, kEllipsoidParamterizationSize, 1>;
template <typename NumType>
struct EllipsoidState {
  Pose3DYawOnly<NumType> pose_;
  ObjectDim<NumType> dimensions_;
};
}  // namespace vslam_types_refactor
