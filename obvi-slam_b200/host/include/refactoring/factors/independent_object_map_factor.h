// obvi-ba-b200 host side.  Stands in for the reference header of the same path
// (include/refactoring/factors/independent_object_map_factor.h: factory createIndependentObjectMapFactor at :35-40).
// Long-term-map prior on one ellipsoid, r = (Sigma^-1)^(1/2) (e - e_map); the information square root is formed by the
// backend.  The object carries the map estimate as the raw 7-vector (x y z yaw dx dy dz,
// vslam_obj_opt_types_refactor.h:132-152) and the 7x7 covariance, row-major, to obvi_factor_add_ltm_prior.
#pragma once

#include <ceres/autodiff_cost_function.h>
#include <refactoring/types/vslam_basic_types_refactor.h>
#include <refactoring/types/vslam_obj_opt_types_refactor.h>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class IndependentObjectMapFactor {
  static constexpr int kDim = kEllipsoidParamterizationSize;
  double map_estimate_[7];
  double sigma_[49];

 public:
  using Cost = ceres::AutoDiffCostFunction<IndependentObjectMapFactor, kDim, kDim>;

  IndependentObjectMapFactor(const EllipsoidState<double>& e_map, const Covariance<double, kDim>& sigma) {
    static_assert(kDim == 7, "yaw-only ellipsoid parameterisation expected");
    for (int a = 0; a < 3; a++) {
      map_estimate_[a] = e_map.pose_.transl_(a);
      map_estimate_[4 + a] = e_map.dimensions_(a);
    }
    map_estimate_[3] = e_map.pose_.yaw_;
    obvi_shim::copySquare<7>(sigma, sigma_);
  }

  static Cost* createIndependentObjectMapFactor(const EllipsoidState<double>& e_map, const Covariance<double, kDim>& sigma) {
    return new Cost(new IndependentObjectMapFactor(e_map, sigma));
  }

  // called by the shim's Problem::AddResidualBlock with the parameter blocks the caller passed
  int obviAdd(obvi_problem* problem, double* const* blocks, double huber, obvi_factor_id* id) const {
    return obvi_factor_add_ltm_prior(problem, blocks[0], map_estimate_, sigma_, huber, id);
  }
};

}  // namespace vslam_types_refactor
