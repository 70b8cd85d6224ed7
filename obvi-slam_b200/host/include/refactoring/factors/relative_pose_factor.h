// Drop-in replacement of include/refactoring/factors/relative_pose_factor.h (createRelativePoseFactor, :74-76).
// A NaN information matrix makes obvi_factor_add_rel_pose return OBVI_ERR_NUMERIC (the reference exits there,
// src/refactoring/factors/relative_pose_factor.cpp:14-18); ceres::Problem::AddResidualBlock turns that into an exception.
#ifndef UT_VSLAM_RELATIVE_POSE_FACTOR_H
#define UT_VSLAM_RELATIVE_POSE_FACTOR_H

#include <ceres/autodiff_cost_function.h>
#include <refactoring/types/vslam_basic_types_refactor.h>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class RelativePoseFactor {
 public:
  RelativePoseFactor(const Pose3D<double>& measured_pose_deviation, const Covariance<double, 6>& pose_deviation_cov) {
    const auto R = measured_pose_deviation.orientation_.toRotationMatrix();
    for (int i = 0; i < 3; i++) {
      t_[i] = measured_pose_deviation.transl_(i);
      for (int j = 0; j < 3; j++) R_[3 * i + j] = R(i, j);
    }
    obvi_shim::copySquare<6>(pose_deviation_cov, cov_);
  }
  // parameter order (pose before, pose after)
  int obviAdd(obvi_problem* p, double* const* blocks, double huber, obvi_factor_id* id) const {
    return obvi_factor_add_rel_pose(p, blocks[0], blocks[1], t_, R_, cov_, huber, id);
  }
  static ceres::AutoDiffCostFunction<RelativePoseFactor, 6, 6, 6>* createRelativePoseFactor(const Pose3D<double>& measured_pose_deviation,
                                                                                             const Covariance<double, 6>& pose_deviation_cov) {
    return new ceres::AutoDiffCostFunction<RelativePoseFactor, 6, 6, 6>(new RelativePoseFactor(measured_pose_deviation, pose_deviation_cov));
  }

 private:
  double t_[3];
  double R_[9];
  double cov_[36];
};
}  // namespace vslam_types_refactor
#endif  // UT_VSLAM_RELATIVE_POSE_FACTOR_H
