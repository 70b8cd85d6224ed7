// Drop-in replacement of include/refactoring/factors/independent_object_map_factor.h
// (createIndependentObjectMapFactor, :35-40): long-term-map prior r = (Sigma^-1)^(1/2) (e - e_map).
#ifndef UT_VSLAM_INDEPENDENT_OBJECT_MAP_FACTOR_H
#define UT_VSLAM_INDEPENDENT_OBJECT_MAP_FACTOR_H

#include <ceres/autodiff_cost_function.h>
#include <refactoring/types/vslam_basic_types_refactor.h>
#include <refactoring/types/vslam_obj_opt_types_refactor.h>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class IndependentObjectMapFactor {
 public:
  IndependentObjectMapFactor(const EllipsoidState<double>& ellipsoid_mean, const Covariance<double, kEllipsoidParamterizationSize>& covariance) {
    // raw ellipsoid layout (x y z yaw dx dy dz), vslam_obj_opt_types_refactor.h:132-152 (CONSTRAIN_ELLIPSOID_ORIENTATION)
    for (int i = 0; i < 3; i++) { mean_[i] = ellipsoid_mean.pose_.transl_(i); mean_[4 + i] = ellipsoid_mean.dimensions_(i); }
    mean_[3] = ellipsoid_mean.pose_.yaw_;
    obvi_shim::copySquare<7>(covariance, cov_);
  }
  int obviAdd(obvi_problem* p, double* const* blocks, double huber, obvi_factor_id* id) const {
    return obvi_factor_add_ltm_prior(p, blocks[0], mean_, cov_, huber, id);
  }
  static ceres::AutoDiffCostFunction<IndependentObjectMapFactor, kEllipsoidParamterizationSize, kEllipsoidParamterizationSize>*
  createIndependentObjectMapFactor(const EllipsoidState<double>& ellipsoid_mean,
                                   const Covariance<double, kEllipsoidParamterizationSize>& covariance) {
    return new ceres::AutoDiffCostFunction<IndependentObjectMapFactor, kEllipsoidParamterizationSize, kEllipsoidParamterizationSize>(
        new IndependentObjectMapFactor(ellipsoid_mean, covariance));
  }

 private:
  double mean_[7];
  double cov_[49];
};
}  // namespace vslam_types_refactor
#endif  // UT_VSLAM_INDEPENDENT_OBJECT_MAP_FACTOR_H
