// Drop-in replacement of include/refactoring/factors/parameter_prior.h (createParameterPrior<N>, :36-45):
// r = (x[idx] - mean) / std_dev, added by the reference without a loss function.
#ifndef UT_VSLAM_PARAMETER_PRIOR_H
#define UT_VSLAM_PARAMETER_PRIOR_H

#include <ceres/autodiff_cost_function.h>

#include <cstddef>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class ParameterPrior {
 public:
  ParameterPrior(const size_t& param_idx, const double& param_mean, const double& param_std_dev)
      : param_idx_(param_idx), param_mean_(param_mean), param_std_dev_(param_std_dev) {}
  int obviAdd(obvi_problem* p, double* const* blocks, double huber, obvi_factor_id* id) const {
    return obvi_factor_add_param_prior(p, blocks[0], (int)param_idx_, param_mean_, param_std_dev_, huber, id);
  }
  template <int ParamBlockSize>
  static ceres::AutoDiffCostFunction<ParameterPrior, 1, ParamBlockSize>* createParameterPrior(const size_t& param_idx, const double& param_mean,
                                                                                             const double& param_std_dev) {
    return new ceres::AutoDiffCostFunction<ParameterPrior, 1, ParamBlockSize>(new ParameterPrior(param_idx, param_mean, param_std_dev));
  }

 private:
  size_t param_idx_;
  double param_mean_;
  double param_std_dev_;
};
}  // namespace vslam_types_refactor
#endif  // UT_VSLAM_PARAMETER_PRIOR_H
