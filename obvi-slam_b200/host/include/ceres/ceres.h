// Header-only Ceres-compatibility shim over the obvi_ba C ABI (include/obvi_ba.h).
//
// It reproduces, in namespace ceres, exactly the API subset the reference's bundle-adjustment path uses
// (grep over the reference's include/ and src/):
//   Problem::{AddParameterBlock, AddResidualBlock, RemoveResidualBlock, RemoveParameterBlock, SetParameterBlockConstant,
//             SetParameterBlockVariable, IsParameterBlockConstant, GetResidualBlocks, GetParameterBlocksForResidualBlock,
//             NumResidualBlocks, NumParameterBlocks, Evaluate}
//   Solver::Options (the fields object_pose_graph_optimizer.h:651-672 sets), Solver::Summary, IterationSummary, Solve,
//   CostFunction, AutoDiffCostFunction, SizedCostFunction, LossFunction, HuberLoss, IterationCallback, ResidualBlockId,
//   Covariance::{Options, Compute, GetCovarianceBlock} on ellipsoid blocks (long-term-map extraction).
// so that include/refactoring/optimization/{residual_creator.h, object_pose_graph_optimizer.h},
// include/refactoring/offline/offline_problem_runner.h and include/run_optimization_utils/* compile against it unchanged
// once the factor headers next to this file (refactoring/factors/*.h) replace the reference's.
// Ownership follows Ceres' defaults: the Problem deletes the cost and loss objects it is given.
#ifndef OBVI_CERES_SHIM_CERES_H_
#define OBVI_CERES_SHIM_CERES_H_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "obvi_ba.h"

namespace ceres {

// ---------------------------------------------------------------------------------------------- enums
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE = 0, NO_CONVERGENCE = 1, FAILURE = 2, USER_SUCCESS = 3, USER_FAILURE = 4 };
enum CallbackReturnType { SOLVER_CONTINUE, SOLVER_ABORT, SOLVER_TERMINATE_SUCCESSFULLY };
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };

// ---------------------------------------------------------------------------------------------- loss functions
class LossFunction {
 public:
  virtual ~LossFunction() {}
  // the Huber parameter a (<= 0: trivial loss); the only loss family the reference uses
  virtual double obvi_huber_parameter() const = 0;
};
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a) {}
  double obvi_huber_parameter() const override { return a_; }
 private:
  double a_;
};

// ---------------------------------------------------------------------------------------------- cost functions
// A cost function here is a *description* of one of the reference's factors: it knows how to register itself with
// the backend (the residual / Jacobian arithmetic runs in CUDA kernels, not through Evaluate()).
class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual int obvi_add(obvi_problem* p, double* const* parameter_blocks, double huber, obvi_factor_id* id) const = 0;
  int num_residuals() const { return num_residuals_; }
  const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
 protected:
  int num_residuals_ = 0;
  std::vector<int32_t> parameter_block_sizes_;
};

// AutoDiffCostFunction<Functor, kNumResiduals, N0[, N1]>: the functor must provide
//   int obviAdd(obvi_problem*, double* const* blocks, double huber, obvi_factor_id*) const
// (the replacement factor headers do).  Takes ownership of the functor, as Ceres does.
template <typename Functor, int kNumResiduals, int N0, int N1 = 0>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* functor) : functor_(functor) {
    num_residuals_ = kNumResiduals;
    parameter_block_sizes_.push_back(N0);
    if (N1 > 0) parameter_block_sizes_.push_back(N1);
  }
  int obvi_add(obvi_problem* p, double* const* blocks, double huber, obvi_factor_id* id) const override {
    return functor_->obviAdd(p, blocks, huber, id);
  }
  const Functor& functor() const { return *functor_; }
 private:
  std::unique_ptr<Functor> functor_;
};

// SizedCostFunction<kNumResiduals, N0[, N1]>: base of the (disabled) symforce reprojection class
// (reprojection_cost_functor_analytic_jacobian.h:18); derived classes override obvi_add.
template <int kNumResiduals, int N0, int N1 = 0>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() {
    num_residuals_ = kNumResiduals;
    parameter_block_sizes_.push_back(N0);
    if (N1 > 0) parameter_block_sizes_.push_back(N1);
  }
};

// ---------------------------------------------------------------------------------------------- ids
struct ResidualBlock {
  obvi_factor_id id;
  std::vector<double*> parameter_blocks;
  CostFunction* cost;
  LossFunction* loss;
};
typedef ResidualBlock* ResidualBlockId;

struct CRSMatrix { int num_rows = 0, num_cols = 0; std::vector<int> cols, rows; std::vector<double> values; };

// ---------------------------------------------------------------------------------------------- problem
class Problem {
 public:
  struct Options { Ownership cost_function_ownership = TAKE_OWNERSHIP; Ownership loss_function_ownership = TAKE_OWNERSHIP; bool enable_fast_removal = false; };
  struct EvaluateOptions {
    std::vector<double*> parameter_blocks;
    std::vector<ResidualBlockId> residual_blocks;
    bool apply_loss_function = true;
    int num_threads = 1;
  };

  Problem() { init(); }
  explicit Problem(const Options& o) : options_(o) { init(); }
  Problem(const Problem&) = delete;
  Problem& operator=(const Problem&) = delete;
  ~Problem() {
    for (auto& kv : blocks_) release(kv.second.get());
    if (handle_) obvi_problem_destroy(handle_);
  }

  void AddParameterBlock(double* values, int size) {
    check(obvi_param_add(handle_, values, size), "AddParameterBlock");
    param_sizes_[values] = size;
  }
  ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0) {
    double* blocks[1] = {x0};
    return add(cost, loss, blocks, 1);
  }
  ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1) {
    double* blocks[2] = {x0, x1};
    return add(cost, loss, blocks, 2);
  }
  ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, const std::vector<double*>& x) {
    return add(cost, loss, x.data(), (int)x.size());
  }
  void RemoveResidualBlock(ResidualBlockId rb) {
    check(obvi_factor_remove(handle_, rb->id), "RemoveResidualBlock");
    release(rb);
    blocks_.erase(rb->id);
  }
  void RemoveParameterBlock(double* values) {
    check(obvi_param_remove(handle_, values), "RemoveParameterBlock");
    for (auto it = blocks_.begin(); it != blocks_.end();) {  // Ceres removes the dependent residual blocks too
      bool uses = false;
      for (double* p : it->second->parameter_blocks) uses |= (p == values);
      if (uses) { release(it->second.get()); it = blocks_.erase(it); } else { ++it; }
    }
    param_sizes_.erase(values);
  }
  void SetParameterBlockConstant(double* values) { check(obvi_param_set_constant(handle_, values, 1), "SetParameterBlockConstant"); }
  void SetParameterBlockVariable(double* values) { check(obvi_param_set_constant(handle_, values, 0), "SetParameterBlockVariable"); }
  bool IsParameterBlockConstant(const double* values) const {
    int c = 0;
    check(obvi_param_is_constant(handle_, values, &c), "IsParameterBlockConstant");
    return c != 0;
  }
  bool HasParameterBlock(const double* values) const { int c; return obvi_param_is_constant(handle_, values, &c) == OBVI_OK; }
  int NumResidualBlocks() const { return (int)obvi_num_factors(handle_); }
  int NumParameterBlocks() const { return (int)param_sizes_.size(); }
  int ParameterBlockSize(const double* values) const { auto it = param_sizes_.find(values); return it == param_sizes_.end() ? 0 : it->second; }
  void GetResidualBlocks(std::vector<ResidualBlockId>* out) const {
    int64_t n = 0;
    check(obvi_residual_blocks(handle_, nullptr, nullptr, nullptr, 0, &n), "GetResidualBlocks");
    std::vector<obvi_factor_id> ids(n);
    check(obvi_residual_blocks(handle_, ids.data(), nullptr, nullptr, n, &n), "GetResidualBlocks");
    out->clear();
    for (obvi_factor_id id : ids) out->push_back(blocks_.at(id).get());
  }
  void GetParameterBlocks(std::vector<double*>* out) const {
    out->clear();
    for (auto& kv : param_sizes_) out->push_back(const_cast<double*>(kv.first));
  }
  void GetParameterBlocksForResidualBlock(const ResidualBlockId rb, std::vector<double*>* out) const { *out = rb->parameter_blocks; }
  const CostFunction* GetCostFunctionForResidualBlock(const ResidualBlockId rb) const { return rb->cost; }
  const LossFunction* GetLossFunctionForResidualBlock(const ResidualBlockId rb) const { return rb->loss; }

  // Problem::Evaluate as the reference calls it: all residual blocks with apply_loss_function = false for the raw
  // residuals (object_pose_graph_optimizer.h:679-693), and with explicit residual / parameter block lists plus a CRS
  // Jacobian for the long-term-map rank analysis (long_term_object_map_extraction.cpp:251-252,591-598).
  bool Evaluate(const EvaluateOptions& o, double* cost, std::vector<double>* residuals, std::vector<double>* gradient,
                CRSMatrix* jacobian) {
    const int loss = o.apply_loss_function ? 1 : 0;
    const bool subset = !o.residual_blocks.empty();
    if ((cost || residuals) && subset && (int)o.residual_blocks.size() != NumResidualBlocks()) { last_error_ = "Evaluate: cost / residuals of a residual-block subset are not supported"; return false; }
    if (cost || residuals) {
      int64_t n = 0;
      double c = 0;
      if (!residuals) {
        if (obvi_evaluate(handle_, loss, &c, nullptr, 0, nullptr) != OBVI_OK) { last_error_ = obvi_last_error(handle_); return false; }
      } else {
        // ONE evaluation: the length of the residual vector is known here (sum of the blocks' residual counts), so the sizing
        // call of the C ABI -- a second full linearisation -- is not needed.
        // obvi_evaluate concatenates in order of addition; reorder when the caller lists the blocks differently
        {
          std::vector<ResidualBlockId> live;
          GetResidualBlocks(&live);
          for (ResidualBlockId rb : live) n += rb->cost->num_residuals();
        }
        std::vector<double> all(n, 0.0);
        int64_t got = 0;
        if (obvi_evaluate(handle_, loss, &c, all.data(), n, &got) != OBVI_OK) { last_error_ = obvi_last_error(handle_); return false; }
        if (got != n) { last_error_ = "Evaluate: residual count mismatch between the shim and the backend"; return false; }
        if (!subset) { *residuals = all; }
        else {
          std::vector<ResidualBlockId> order;
          GetResidualBlocks(&order);
          std::unordered_map<ResidualBlockId, std::pair<int64_t, int>> at;
          int64_t w = 0;
          for (ResidualBlockId rb : order) { at[rb] = {w, rb->cost->num_residuals()}; w += rb->cost->num_residuals(); }
          residuals->clear();
          for (ResidualBlockId rb : o.residual_blocks) { auto& e = at.at(rb); residuals->insert(residuals->end(), all.begin() + e.first, all.begin() + e.first + e.second); }
        }
      }
      if (cost) *cost = c;
    }
    if (gradient || jacobian) {
      std::vector<obvi_factor_id> ids;
      for (ResidualBlockId rb : o.residual_blocks) ids.push_back(rb->id);
      std::vector<double*> pbs(o.parameter_blocks.begin(), o.parameter_blocks.end());
      int64_t nr = 0, nc = 0, nz = 0;
      const obvi_factor_id* idp = ids.empty() ? nullptr : ids.data();
      double* const* pbp = pbs.empty() ? nullptr : pbs.data();
      if (obvi_evaluate_jacobian(handle_, loss, idp, (int64_t)ids.size(), pbp, (int64_t)pbs.size(), &nr, &nc, &nz, nullptr, nullptr, nullptr, nullptr) != OBVI_OK) { last_error_ = obvi_last_error(handle_); return false; }
      std::vector<int32_t> rows(nr + 1), cols(nz);
      std::vector<double> vals(nz), grad(nc);
      if (obvi_evaluate_jacobian(handle_, loss, idp, (int64_t)ids.size(), pbp, (int64_t)pbs.size(), &nr, &nc, &nz, rows.data(), cols.data(), vals.data(), grad.data()) != OBVI_OK) { last_error_ = obvi_last_error(handle_); return false; }
      if (gradient) *gradient = grad;
      if (jacobian) { jacobian->num_rows = (int)nr; jacobian->num_cols = (int)nc; jacobian->rows.assign(rows.begin(), rows.end()); jacobian->cols.assign(cols.begin(), cols.end()); jacobian->values = vals; }
    }
    return true;
  }

  obvi_problem* obvi_handle() const { return handle_; }
  const std::string& obvi_last_error_message() const { return last_error_; }

 private:
  void init() {
    const char* dev = std::getenv("OBVI_CUDA_DEVICE");
    if (obvi_problem_create(dev ? std::atoi(dev) : 0, &handle_) != OBVI_OK)
      throw std::runtime_error(std::string("obvi_problem_create failed: ") + obvi_last_error(nullptr));
  }
  void check(int rc, const char* what) const {
    if (rc != OBVI_OK) throw std::runtime_error(std::string(what) + ": " + obvi_last_error(handle_));
  }
  ResidualBlockId add(CostFunction* cost, LossFunction* loss, double* const* x, int n) {
    if ((int)cost->parameter_block_sizes().size() != n) throw std::runtime_error("AddResidualBlock: wrong number of parameter blocks");
    for (int i = 0; i < n; i++) {
      AddParameterBlock(x[i], cost->parameter_block_sizes()[i]);
    }
    obvi_factor_id id = 0;
    check(cost->obvi_add(handle_, x, loss ? loss->obvi_huber_parameter() : 0.0, &id), "AddResidualBlock");
    cost_refs_[cost]++;
    if (loss) loss_refs_[loss]++;
    std::unique_ptr<ResidualBlock> rb(new ResidualBlock{id, std::vector<double*>(x, x + n), cost, loss});
    ResidualBlockId out = rb.get();
    blocks_[id] = std::move(rb);
    return out;
  }
  // Ceres reference-counts the cost / loss objects it owns: one LossFunction (or CostFunction) may be shared by many residual
  // blocks and is deleted when the last of them goes.
  void release(ResidualBlock* rb) {
    if (rb->cost && options_.cost_function_ownership == TAKE_OWNERSHIP && --cost_refs_[rb->cost] == 0) { cost_refs_.erase(rb->cost); delete rb->cost; }
    if (rb->loss && options_.loss_function_ownership == TAKE_OWNERSHIP && --loss_refs_[rb->loss] == 0) { loss_refs_.erase(rb->loss); delete rb->loss; }
    rb->cost = nullptr; rb->loss = nullptr;
  }
  std::unordered_map<const CostFunction*, int> cost_refs_;
  std::unordered_map<const LossFunction*, int> loss_refs_;
  Options options_;
  obvi_problem* handle_ = nullptr;
  std::unordered_map<obvi_factor_id, std::unique_ptr<ResidualBlock>> blocks_;
  std::unordered_map<const double*, int> param_sizes_;
  std::string last_error_;
};

// ---------------------------------------------------------------------------------------------- solver
struct IterationSummary {
  int iteration = 0;
  bool step_is_valid = false, step_is_nonmonotonic = false, step_is_successful = false;
  double cost = 0, cost_change = 0, gradient_max_norm = 0, gradient_norm = 0, step_norm = 0, relative_decrease = 0,
         trust_region_radius = 0, eta = 0, step_size = 0;
  int line_search_function_evaluations = 0, linear_solver_iterations = 0;
  double iteration_time_in_seconds = 0, step_solver_time_in_seconds = 0, cumulative_time_in_seconds = 0;
};

class IterationCallback {
 public:
  virtual ~IterationCallback() {}
  virtual CallbackReturnType operator()(const IterationSummary& summary) = 0;
};

class Solver {
 public:
  struct Options {
    // fields the reference sets (object_pose_graph_optimizer.h:651-672)
    std::vector<IterationCallback*> callbacks;
    bool update_state_every_iteration = false;
    int max_num_iterations = 50;
    int num_threads = 1;                                 // accepted, unused (the GPU is the thread pool)
    LinearSolverType linear_solver_type = SPARSE_SCHUR;  // accepted; the backend always eliminates points and objects
    bool use_nonmonotonic_steps = false;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16;
    // Ceres defaults the reference leaves alone
    double min_trust_region_radius = 1e-32, min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    int max_consecutive_nonmonotonic_steps = 5, max_num_consecutive_invalid_steps = 5;
    bool minimizer_progress_to_stdout = false;
  };
  struct Summary {
    TerminationType termination_type = FAILURE;
    std::string message;
    double initial_cost = 0, final_cost = 0, fixed_cost = 0;
    std::vector<IterationSummary> iterations;
    int num_successful_steps = 0, num_unsuccessful_steps = 0;
    double preprocessor_time_in_seconds = 0, minimizer_time_in_seconds = 0, total_time_in_seconds = 0,
           linear_solver_time_in_seconds = 0, residual_evaluation_time_in_seconds = 0, jacobian_evaluation_time_in_seconds = 0;
    int num_parameter_blocks_reduced = 0, num_parameters_reduced = 0, num_residual_blocks_reduced = 0, num_residuals_reduced = 0;
    int num_threads_used = 1;
    int64_t obvi_kernel_launches = 0, obvi_pcg_iterations = 0;
    bool IsSolutionUsable() const { return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE || termination_type == USER_SUCCESS; }
    std::string BriefReport() const {
      std::ostringstream s;
      s << "obvi_ba Report: Iterations: " << iterations.size() << ", Initial cost: " << initial_cost << ", Final cost: " << final_cost
        << ", Termination: " << (termination_type == CONVERGENCE ? "CONVERGENCE" : termination_type == NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE");
      return s.str();
    }
    std::string FullReport() const {
      std::ostringstream s;
      s << BriefReport() << "\n" << message << "\nReduced parameters " << num_parameters_reduced << ", residuals " << num_residuals_reduced
        << "\nTime (s): total " << total_time_in_seconds << ", jacobian " << jacobian_evaluation_time_in_seconds << ", linear solver "
        << linear_solver_time_in_seconds << ", residual " << residual_evaluation_time_in_seconds << "\nGPU kernel launches "
        << obvi_kernel_launches << ", PCG iterations " << obvi_pcg_iterations << "\n";
      return s.str();
    }
  };
};

inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  obvi_solver_options o;
  obvi_solver_options_init(&o);
  o.max_num_iterations = options.max_num_iterations;
  o.use_nonmonotonic_steps = options.use_nonmonotonic_steps ? 1 : 0;
  o.function_tolerance = options.function_tolerance; o.gradient_tolerance = options.gradient_tolerance;
  o.parameter_tolerance = options.parameter_tolerance; o.initial_trust_region_radius = options.initial_trust_region_radius;
  o.max_trust_region_radius = options.max_trust_region_radius; o.min_trust_region_radius = options.min_trust_region_radius;
  o.min_relative_decrease = options.min_relative_decrease; o.min_lm_diagonal = options.min_lm_diagonal; o.max_lm_diagonal = options.max_lm_diagonal;
  o.max_consecutive_nonmonotonic_steps = options.max_consecutive_nonmonotonic_steps;
  o.max_num_consecutive_invalid_steps = options.max_num_consecutive_invalid_steps;
  // Iteration callbacks run on this thread after every iteration summary, in order; SOLVER_ABORT / SOLVER_TERMINATE_SUCCESSFULLY
  // stop the solve there (USER_FAILURE / USER_SUCCESS), and with update_state_every_iteration the parameter blocks hold the
  // current iterate when they run (object_pose_graph_optimizer.h:651-659).
  struct Trampoline {
    const Solver::Options* opt;
    static int32_t call(void* user, const void* it_) {
      const Solver::Options& op = *static_cast<Trampoline*>(user)->opt;
      const obvi_iteration_summary& in = *static_cast<const obvi_iteration_summary*>(it_);
      IterationSummary it;
      it.iteration = in.iteration; it.step_is_valid = in.step_is_valid != 0; it.step_is_successful = in.step_is_successful != 0;
      it.cost = in.cost; it.cost_change = in.cost_change; it.gradient_max_norm = in.gradient_max_norm; it.step_norm = in.step_norm;
      it.relative_decrease = in.relative_decrease; it.trust_region_radius = in.trust_region_radius; it.linear_solver_iterations = in.linear_solver_iterations;
      for (IterationCallback* cb : op.callbacks) {
        const CallbackReturnType r = (*cb)(it);
        if (r == SOLVER_ABORT) return 1;
        if (r == SOLVER_TERMINATE_SUCCESSFULLY) return 2;
      }
      return 0;
    }
  } tramp{&options};
  if (!options.callbacks.empty()) {
    o.iteration_callback = &Trampoline::call; o.iteration_callback_user = &tramp;
    o.update_state_every_iteration = options.update_state_every_iteration ? 1 : 0;
  }
  obvi_summary s;
  std::vector<obvi_iteration_summary> its(options.max_num_iterations + 2);
  *summary = Solver::Summary();
  const int rc = obvi_solve(problem->obvi_handle(), &o, &s, its.data(), (int32_t)its.size());
  if (rc != OBVI_OK) {
    summary->termination_type = FAILURE;
    summary->message = std::string("obvi_solve failed: ") + obvi_last_error(problem->obvi_handle());
    return;
  }
  summary->termination_type = (TerminationType)s.termination_type;
  summary->initial_cost = s.initial_cost; summary->final_cost = s.final_cost; summary->fixed_cost = s.fixed_cost;
  summary->num_successful_steps = s.num_successful_steps; summary->num_unsuccessful_steps = s.num_unsuccessful_steps;
  summary->preprocessor_time_in_seconds = s.preprocessor_time_in_seconds; summary->minimizer_time_in_seconds = s.minimizer_device_time_in_seconds;
  summary->total_time_in_seconds = s.total_time_in_seconds; summary->linear_solver_time_in_seconds = s.linear_solver_time_in_seconds;
  summary->residual_evaluation_time_in_seconds = s.residual_evaluation_time_in_seconds;
  summary->jacobian_evaluation_time_in_seconds = s.jacobian_evaluation_time_in_seconds;
  summary->num_parameter_blocks_reduced = s.num_parameter_blocks_reduced; summary->num_parameters_reduced = s.num_parameters_reduced;
  summary->num_residual_blocks_reduced = s.num_residual_blocks_reduced; summary->num_residuals_reduced = s.num_residuals_reduced;
  summary->obvi_kernel_launches = s.kernel_launches; summary->obvi_pcg_iterations = s.pcg_iterations_total;
  const int n = s.num_iterations < (int)its.size() ? s.num_iterations : (int)its.size();
  for (int i = 0; i < n; i++) {
    IterationSummary it;
    it.iteration = its[i].iteration; it.step_is_valid = its[i].step_is_valid != 0; it.step_is_successful = its[i].step_is_successful != 0;
    it.cost = its[i].cost; it.cost_change = its[i].cost_change; it.gradient_max_norm = its[i].gradient_max_norm; it.step_norm = its[i].step_norm;
    it.relative_decrease = its[i].relative_decrease; it.trust_region_radius = its[i].trust_region_radius;
    it.linear_solver_iterations = its[i].linear_solver_iterations;
    summary->iterations.push_back(it);
  }
  summary->message = summary->termination_type == CONVERGENCE ? "Convergence" : summary->termination_type == NO_CONVERGENCE ? "Maximum number of iterations reached"
                     : summary->termination_type == USER_SUCCESS ? "User callback returned SOLVER_TERMINATE_SUCCESSFULLY"
                     : summary->termination_type == USER_FAILURE ? "User callback returned SOLVER_ABORT" : "Failure";
}

// ---------------------------------------------------------------------------------------------- covariance
// ceres::Covariance as the long-term-map extraction uses it (long_term_object_map_extraction.cpp:412-440,
// long_term_object_map_extraction.h:269-345,459-520): Compute() on a list of (ellipsoid, ellipsoid) block pairs, then
// GetCovarianceBlock() per pair (7x7 row-major).  Only ellipsoid blocks are supported.
enum CovarianceAlgorithmType { DENSE_SVD, SPARSE_QR };
enum SparseLinearAlgebraLibraryType { SUITE_SPARSE, EIGEN_SPARSE, NO_SPARSE };
class Covariance {
 public:
  struct Options {
    int num_threads = 1;
    CovarianceAlgorithmType algorithm_type = SPARSE_QR;
    SparseLinearAlgebraLibraryType sparse_linear_algebra_library_type = SUITE_SPARSE;
    bool apply_loss_function = true;
  };
  explicit Covariance(const Options& o) : options_(o) {}
  bool Compute(const std::vector<std::pair<const double*, const double*>>& blocks, Problem* problem) {
    blocks_.clear(); values_.clear();
    if (!options_.apply_loss_function) return false;   // the backend evaluates the loss-corrected Jacobian (Ceres' default)
    std::vector<double*> a, b;
    for (auto& pr : blocks) {
      if (problem->ParameterBlockSize(pr.first) != 7 || problem->ParameterBlockSize(pr.second) != 7) return false;
      a.push_back(const_cast<double*>(pr.first)); b.push_back(const_cast<double*>(pr.second));
    }
    values_.assign(blocks.size() * 49, 0.0);
    if (obvi_object_covariances(problem->obvi_handle(), (int64_t)blocks.size(), a.data(), b.data(), values_.data()) != OBVI_OK) { values_.clear(); return false; }
    for (size_t i = 0; i < blocks.size(); i++) blocks_[blocks[i]] = i;
    return true;
  }
  bool GetCovarianceBlock(const double* p1, const double* p2, double* out) const {
    auto it = blocks_.find(std::make_pair(p1, p2));
    if (it != blocks_.end()) { for (int i = 0; i < 49; i++) out[i] = values_[it->second * 49 + i]; return true; }
    it = blocks_.find(std::make_pair(p2, p1));   // Ceres serves the transposed block as well
    if (it == blocks_.end()) return false;
    for (int r = 0; r < 7; r++) for (int c = 0; c < 7; c++) out[r * 7 + c] = values_[it->second * 49 + c * 7 + r];
    return true;
  }
 private:
  struct PairHash { size_t operator()(const std::pair<const double*, const double*>& p) const { return std::hash<const double*>()(p.first) * 1000003u ^ std::hash<const double*>()(p.second); } };
  Options options_;
  std::unordered_map<std::pair<const double*, const double*>, size_t, PairHash> blocks_;
  std::vector<double> values_;
};

}  // namespace ceres
#endif  // OBVI_CERES_SHIM_CERES_H_
