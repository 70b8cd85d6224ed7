// Drop-in replacement of include/refactoring/factors/shape_prior_factor.h (createShapeDimPrior, :76-80).
#ifndef UT_VSLAM_SHAPE_PRIOR_FACTOR_H
#define UT_VSLAM_SHAPE_PRIOR_FACTOR_H

#include <ceres/autodiff_cost_function.h>
#include <refactoring/types/vslam_basic_types_refactor.h>
#include <refactoring/types/vslam_obj_opt_types_refactor.h>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class ShapePriorFactor {
 public:
  ShapePriorFactor(const ObjectDim<double>& shape_dim_mean, const Covariance<double, 3>& shape_dim_covariance) {
    for (int i = 0; i < 3; i++) mean_[i] = shape_dim_mean(i);
    obvi_shim::copySquare<3>(shape_dim_covariance, cov_);
  }
  int obviAdd(obvi_problem* p, double* const* blocks, double huber, obvi_factor_id* id) const {
    return obvi_factor_add_shape_prior(p, blocks[0], mean_, cov_, huber, id);
  }
  static ceres::AutoDiffCostFunction<ShapePriorFactor, 3, kEllipsoidParamterizationSize>* createShapeDimPrior(
      const vslam_types_refactor::ObjectDim<double>& dimension_prior_mean, const vslam_types_refactor::Covariance<double, 3>& dimension_cov) {
    return new ceres::AutoDiffCostFunction<ShapePriorFactor, 3, kEllipsoidParamterizationSize>(new ShapePriorFactor(dimension_prior_mean, dimension_cov));
  }

 private:
  double mean_[3];
  double cov_[9];
};
}  // namespace vslam_types_refactor
#endif  // UT_VSLAM_SHAPE_PRIOR_FACTOR_H
