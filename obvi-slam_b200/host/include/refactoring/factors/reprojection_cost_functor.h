// Drop-in replacement of include/refactoring/factors/reprojection_cost_functor.h (class name and static factory
// signature identical, reprojection_cost_functor.h:135-144).  The functor no longer carries templated arithmetic for
// Ceres autodiff: it records the factory arguments and registers a reprojection factor with the CUDA backend, which
// evaluates the same residual (csrc/factors.cuh: reproj_residual_jacobian) with a closed-form Jacobian.
#ifndef UT_VSLAM_REPROJECTION_COST_FUNCTOR_H
#define UT_VSLAM_REPROJECTION_COST_FUNCTOR_H

#include <ceres/autodiff_cost_function.h>
#include <refactoring/types/vslam_basic_types_refactor.h>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class ReprojectionCostFunctor {
 public:
  ReprojectionCostFunctor(const PixelCoord<double>& image_feature, const CameraIntrinsicsMat<double>& intrinsics,
                          const CameraExtrinsics<double>& extrinsics, const double& reprojection_error_std_dev)
      : camera_(obvi_shim::makeCamera(intrinsics, extrinsics)), sigma_(reprojection_error_std_dev) {
    pixel_[0] = image_feature(0);
    pixel_[1] = image_feature(1);
  }

  // parameter order (pose, point), residual_creator.h:263-264
  int obviAdd(obvi_problem* p, double* const* blocks, double huber, obvi_factor_id* id) const {
    int cam = -1;
    const int rc = obvi_shim::registerCamera(p, camera_, &cam);
    if (rc != OBVI_OK) return rc;
    return obvi_factor_add_reproj(p, blocks[0], blocks[1], cam, pixel_, sigma_, huber, id);
  }

  static ceres::AutoDiffCostFunction<ReprojectionCostFunctor, 2, 6, 3>* create(
      const vslam_types_refactor::CameraIntrinsicsMat<double>& intrinsics,
      const vslam_types_refactor::CameraExtrinsics<double>& extrinsics,
      const vslam_types_refactor::PixelCoord<double>& feature_pixel, const double& reprojection_error_std_dev) {
    ReprojectionCostFunctor* residual = new ReprojectionCostFunctor(feature_pixel, intrinsics, extrinsics, reprojection_error_std_dev);
    return new ceres::AutoDiffCostFunction<ReprojectionCostFunctor, 2, 6, 3>(residual);
  }

 private:
  obvi_shim::CameraData camera_;
  double pixel_[2];
  double sigma_;
};
}  // namespace vslam_types_refactor
#endif  // UT_VSLAM_REPROJECTION_COST_FUNCTOR_H
