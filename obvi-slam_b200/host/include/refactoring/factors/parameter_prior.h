// obvi-ba-b200 host side.  Stands in for the reference header of the same path (include/refactoring/factors/parameter_prior.h:
// class at :19-51, factory createParameterPrior<N> at :36-45).  One scalar residual (x[k] - mu) / sigma on coordinate k of an
// N-wide block; the long-term-map rank repair adds it without a loss function (long_term_object_map_extraction.cpp:817, 869,
// 916).  Nothing is evaluated here: the object only carries the three constants to obvi_factor_add_param_prior.
#pragma once

#include <ceres/autodiff_cost_function.h>

#include <cstddef>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class ParameterPrior {
  struct Constants { int coordinate; double mu, sigma; } c_;

 public:
  template <int N>
  using Cost = ceres::AutoDiffCostFunction<ParameterPrior, 1, N>;

  ParameterPrior(const size_t& k, const double& mu, const double& sigma) : c_{static_cast<int>(k), mu, sigma} {}

  template <int N>
  static Cost<N>* createParameterPrior(const size_t& k, const double& mu, const double& sigma) {
    return new Cost<N>(new ParameterPrior(k, mu, sigma));
  }

  // called by the shim's Problem::AddResidualBlock with the parameter blocks the caller passed
  int obviAdd(obvi_problem* problem, double* const* blocks, double huber, obvi_factor_id* id) const {
    return obvi_factor_add_param_prior(problem, blocks[0], c_.coordinate, c_.mu, c_.sigma, huber, id);
  }
};

}  // namespace vslam_types_refactor
