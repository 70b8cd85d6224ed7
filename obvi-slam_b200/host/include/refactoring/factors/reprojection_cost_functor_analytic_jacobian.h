// Drop-in replacement of include/refactoring/factors/reprojection_cost_functor_analytic_jacobian.h (the symforce-generated
// closed-form variant, class name and constructor signature identical, :17-36; disabled in the reference's
// residual_creator.h but kept selectable there).  It registers the same reprojection factor with the CUDA backend as
// ReprojectionCostFunctor: the backend's Jacobian is already closed-form.  The generated code's regularisation
// (epsilon = 1e-15 inside sqrt(|omega|^2 + eps) and max(z, eps), :66,160,591) is NOT reproduced: away from the two
// singularities (zero rotation, point in the camera plane) the two variants agree to rounding, at them the backend
// follows the autodiff functor's branches (vslam_math_util.h:121-141).
#ifndef UT_VSLAM_REFACTORING_REPROJECTION_COST_FUNCTOR_ANALYTIC_JACOBIAN_H
#define UT_VSLAM_REFACTORING_REPROJECTION_COST_FUNCTOR_ANALYTIC_JACOBIAN_H

#include <ceres/sized_cost_function.h>
#include <refactoring/types/vslam_basic_types_refactor.h>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class ReprojectionCostFunctorAnalyticJacobian : public ceres::SizedCostFunction<2, 6, 3> {
 public:
  ReprojectionCostFunctorAnalyticJacobian(const vslam_types_refactor::PixelCoord<double>& image_feature,
                                          const vslam_types_refactor::CameraIntrinsicsMat<double>& intrinsics,
                                          const vslam_types_refactor::CameraExtrinsics<double>& extrinsics,
                                          const double& reprojection_error_std_dev)
      : camera_(obvi_shim::makeCamera(intrinsics, extrinsics)), sigma_(reprojection_error_std_dev) {
    pixel_[0] = image_feature(0);
    pixel_[1] = image_feature(1);
  }
  virtual ~ReprojectionCostFunctorAnalyticJacobian() = default;

  // parameters[0] = pose, parameters[1] = point (:60-61)
  int obvi_add(obvi_problem* p, double* const* blocks, double huber, obvi_factor_id* id) const override {
    int cam = -1;
    const int rc = obvi_shim::registerCamera(p, camera_, &cam);
    if (rc != OBVI_OK) return rc;
    return obvi_factor_add_reproj(p, blocks[0], blocks[1], cam, pixel_, sigma_, huber, id);
  }

 private:
  obvi_shim::CameraData camera_;
  double pixel_[2];
  double sigma_;
};
}  // namespace vslam_types_refactor
#endif  // UT_VSLAM_REFACTORING_REPROJECTION_COST_FUNCTOR_ANALYTIC_JACOBIAN_H
