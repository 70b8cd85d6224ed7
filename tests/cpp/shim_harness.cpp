// Restates the reference's call patterns against the Ceres-compatibility shim + replacement factor headers
// (residual_creator.h:101,162,251,337; object_pose_graph_optimizer.h:417-472,651-693; offline_problem_runner.h:689-801):
//   AddParameterBlock / AddResidualBlock(F::create(...), new ceres::HuberLoss(a), p0[, p1]) / SetParameterBlockConstant /
//   Solve / GetResidualBlocks / Evaluate(apply_loss_function = false) / RemoveResidualBlock / Solve again.
// Reads a graph dumped by tests/test_shim.py, writes a small result file.  Compile-checked on CPU, run on the GPU box.
#include <ceres/ceres.h>
#include <refactoring/factors/bounding_box_factor.h>
#include <refactoring/factors/independent_object_map_factor.h>
#include <refactoring/factors/parameter_prior.h>
#include <refactoring/factors/relative_pose_factor.h>
#include <refactoring/factors/reprojection_cost_functor.h>
#include <refactoring/factors/reprojection_cost_functor_analytic_jacobian.h>
#include <refactoring/factors/shape_prior_factor.h>

#include <algorithm>
#include <cstdio>
#include <functional>
#include <map>
#include <unordered_map>
#include <vector>

using namespace vslam_types_refactor;

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: shim_harness graph.bin result.bin\n"); return 2; }
  std::vector<double> d;
  {
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::fseek(f, 0, SEEK_END); const long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    d.resize(n / 8);
    if (std::fread(d.data(), 8, d.size(), f) != d.size()) return 2;
    std::fclose(f);
  }
  size_t at = 0;
  auto take = [&](size_t n) { const double* p = &d[at]; at += n; return p; };
  const double* h = take(16);
  const int K = (int)h[0], P = (int)h[1], O = (int)h[2], C = (int)h[3];
  const int n_rp = (int)h[4], n_bb = (int)h[5], n_sh = (int)h[6], n_lt = (int)h[7], n_rl = (int)h[8];
  const double hub_rp = h[9], hub_bb = h[10], invalid = h[11], hub_sh = h[12], hub_lt = h[13], hub_rl = h[14];
  // parameter blocks: one heap allocation per block, as in the reference's pose graph (low_level_feature_pose_graph.h:25-65)
  std::vector<std::unique_ptr<double[]>> poses(K), points(P), objs(O);
  for (int k = 0; k < K; k++) { poses[k].reset(new double[6]); const double* v = take(6); std::copy(v, v + 6, poses[k].get()); }
  for (int k = 0; k < P; k++) { points[k].reset(new double[3]); const double* v = take(3); std::copy(v, v + 3, points[k].get()); }
  for (int k = 0; k < O; k++) { objs[k].reset(new double[7]); const double* v = take(7); std::copy(v, v + 7, objs[k].get()); }
  const double* const_pose = take(K);
  std::vector<CameraIntrinsicsMat<double>> intr(C);
  std::vector<CameraExtrinsics<double>> extr(C);
  for (int c = 0; c < C; c++) { const double* v = take(4); intr[c](0, 0) = v[0]; intr[c](1, 1) = v[1]; intr[c](0, 2) = v[2]; intr[c](1, 2) = v[3]; intr[c](2, 2) = 1.0; }
  for (int c = 0; c < C; c++) { const double* v = take(9); for (int i = 0; i < 9; i++) extr[c].orientation_.R.v[i] = v[i]; }
  for (int c = 0; c < C; c++) { const double* v = take(3); for (int i = 0; i < 3; i++) extr[c].transl_(i) = v[i]; }

  ceres::Problem problem;
  std::unordered_map<ceres::ResidualBlockId, std::pair<int, int>> block_info;  // -> (factor type, index), as the optimizer keeps
  for (int n = 0; n < n_rp; n++) {
    const double* v = take(6);
    PixelCoord<double> px; px(0) = v[3]; px(1) = v[4];
    // every third block goes through the (reference-disabled) symforce class: same factor on the backend
    ceres::CostFunction* cf = (n % 3 == 2) ? static_cast<ceres::CostFunction*>(new ReprojectionCostFunctorAnalyticJacobian(px, intr[(int)v[2]], extr[(int)v[2]], v[5]))
                                           : static_cast<ceres::CostFunction*>(ReprojectionCostFunctor::create(intr[(int)v[2]], extr[(int)v[2]], px, v[5]));
    ceres::ResidualBlockId id = problem.AddResidualBlock(cf, new ceres::HuberLoss(hub_rp), poses[(int)v[0]].get(), points[(int)v[1]].get());
    block_info[id] = {0, n};
  }
  for (int n = 0; n < n_bb; n++) {
    const double* v = take(23);
    BbCorners<double> corners; for (int i = 0; i < 4; i++) corners(i) = v[3 + i];
    Covariance<double, 4> cov; for (int i = 0; i < 16; i++) cov.v[i] = v[7 + i];
    ceres::ResidualBlockId id = problem.AddResidualBlock(
        BoundingBoxFactor::createBoundingBoxFactor(invalid, corners, intr[(int)v[2]], extr[(int)v[2]], cov, (ObjectId)v[0], (FrameId)v[1], (CameraId)v[2]),
        new ceres::HuberLoss(hub_bb), objs[(int)v[0]].get(), poses[(int)v[1]].get());
    block_info[id] = {2, n};
  }
  for (int n = 0; n < n_sh; n++) {
    const double* v = take(13);
    ObjectDim<double> mean; for (int i = 0; i < 3; i++) mean(i) = v[1 + i];
    Covariance<double, 3> cov; for (int i = 0; i < 9; i++) cov.v[i] = v[4 + i];
    block_info[problem.AddResidualBlock(ShapePriorFactor::createShapeDimPrior(mean, cov), new ceres::HuberLoss(hub_sh), objs[(int)v[0]].get())] = {3, n};
  }
  for (int n = 0; n < n_lt; n++) {
    const double* v = take(57);
    EllipsoidState<double> e;
    for (int i = 0; i < 3; i++) { e.pose_.transl_(i) = v[1 + i]; e.dimensions_(i) = v[5 + i]; }
    e.pose_.yaw_ = v[4];
    Covariance<double, 7> cov; for (int i = 0; i < 49; i++) cov.v[i] = v[8 + i];
    block_info[problem.AddResidualBlock(IndependentObjectMapFactor::createIndependentObjectMapFactor(e, cov), new ceres::HuberLoss(hub_lt), objs[(int)v[0]].get())] = {4, n};
  }
  // ONE loss object shared by every relative-pose block: legal in Ceres (the Problem reference-counts what it owns)
  ceres::LossFunction* shared_rel_loss = new ceres::HuberLoss(hub_rl);
  for (int n = 0; n < n_rl; n++) {
    const double* v = take(50);
    Pose3D<double> meas; for (int i = 0; i < 3; i++) meas.transl_(i) = v[2 + i];
    for (int i = 0; i < 9; i++) meas.orientation_.R.v[i] = v[5 + i];
    Covariance<double, 6> cov; for (int i = 0; i < 36; i++) cov.v[i] = v[14 + i];
    block_info[problem.AddResidualBlock(RelativePoseFactor::createRelativePoseFactor(meas, cov), shared_rel_loss, poses[(int)v[0]].get(), poses[(int)v[1]].get())] = {5, n};
  }
  for (int k = 0; k < K; k++) {
    problem.AddParameterBlock(poses[k].get(), 6);  // object_pose_graph_optimizer.h:417-422
    if (const_pose[k] != 0.0) problem.SetParameterBlockConstant(poses[k].get()); else problem.SetParameterBlockVariable(poses[k].get());
  }

  // ---- phase one (object_pose_graph_optimizer.h:651-676)
  ceres::Solver::Options options;
  options.max_num_iterations = (int)h[15];
  options.num_threads = 20;
  options.linear_solver_type = ceres::SPARSE_SCHUR;
  options.use_nonmonotonic_steps = true;
  options.function_tolerance = 1e-6; options.gradient_tolerance = 1e-10; options.parameter_tolerance = 1e-8;
  options.initial_trust_region_radius = 100.0; options.max_trust_region_radius = 1e4;
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  std::printf("%s\n", summary.BriefReport().c_str());
  if (!summary.IsSolutionUsable()) return 1;

  // ---- raw residuals + per-block squared norms (object_pose_graph_optimizer.h:679-693, offline_problem_runner.h:689-749)
  std::vector<ceres::ResidualBlockId> residual_block_ids;
  problem.GetResidualBlocks(&residual_block_ids);
  ceres::Problem::EvaluateOptions eval;
  eval.apply_loss_function = false;
  eval.residual_blocks = residual_block_ids;
  std::vector<double> residuals;
  double raw_cost = 0;
  if (!problem.Evaluate(eval, &raw_cost, &residuals, nullptr, nullptr)) return 1;
  const int sizes[6] = {2, 0, 4, 3, 7, 6};
  std::map<double, ceres::ResidualBlockId, std::greater<double>> reproj_by_err;
  size_t w = 0;
  for (ceres::ResidualBlockId id : residual_block_ids) {
    const int type = block_info.at(id).first;
    double e = 0;
    for (int i = 0; i < sizes[type]; i++) e += residuals[w + i] * residuals[w + i];
    w += sizes[type];
    if (type == 0) reproj_by_err[e] = id;
  }
  // ---- exclude the worst 10 % of the reprojection blocks and solve again (phase two)
  const size_t n_excl = (size_t)(reproj_by_err.size() * 0.1);
  size_t c = 0;
  for (auto it = reproj_by_err.begin(); it != reproj_by_err.end() && c < n_excl; ++it, ++c) problem.RemoveResidualBlock(it->second);
  ceres::Solver::Summary summary2;
  ceres::Solve(options, &problem, &summary2);
  std::printf("%s\n", summary2.BriefReport().c_str());

  std::vector<double> out = {summary.initial_cost, summary.final_cost, (double)summary.iterations.size(), raw_cost, (double)residuals.size(),
                             (double)n_excl, summary2.initial_cost, summary2.final_cost, (double)summary2.iterations.size(),
                             (double)problem.NumResidualBlocks(), (double)summary.num_parameters_reduced};
  // ---- long-term-map extraction pattern (long_term_object_map_extraction.h:459-520): diagonal covariance block of
  //      every ellipsoid that has observations, in object order; 49 values each (zeros when Compute fails)
  {
    ceres::Covariance::Options cov_options;
    cov_options.num_threads = 20; cov_options.algorithm_type = ceres::SPARSE_QR;
    ceres::Covariance covariance(cov_options);
    std::vector<std::pair<const double*, const double*>> cov_blocks;
    for (int k = 0; k < O; k++) if (problem.HasParameterBlock(objs[k].get())) cov_blocks.emplace_back(objs[k].get(), objs[k].get());
    const bool ok = covariance.Compute(cov_blocks, &problem);
    out.push_back(ok ? (double)cov_blocks.size() : -1.0);
    for (auto& pr : cov_blocks) {
      double blk[49] = {0};
      if (ok) covariance.GetCovarianceBlock(pr.first, pr.second, blk);
      out.insert(out.end(), blk, blk + 49);
    }
  }
  for (int k = 0; k < K; k++) out.insert(out.end(), poses[k].get(), poses[k].get() + 6);
  for (int k = 0; k < P; k++) out.insert(out.end(), points[k].get(), points[k].get() + 3);
  for (int k = 0; k < O; k++) out.insert(out.end(), objs[k].get(), objs[k].get() + 7);
  // ---- iteration callback with update_state_every_iteration (object_pose_graph_optimizer.h:651-659): stop after iteration 2
  {
    struct StopAt2 : ceres::IterationCallback {
      const double* watched; std::vector<double> seen; int calls = 0;
      ceres::CallbackReturnType operator()(const ceres::IterationSummary& it) override {
        calls++; seen.push_back(watched[0]);
        return it.iteration >= 2 ? ceres::SOLVER_TERMINATE_SUCCESSFULLY : ceres::SOLVER_CONTINUE;
      }
    } cb;
    // perturb one variable pose so that the first iterations move it visibly
    double* wp = poses[K - 1].get(); wp[0] += 0.05; cb.watched = wp;
    ceres::Solver::Options o3 = options;
    o3.callbacks.push_back(&cb); o3.update_state_every_iteration = true;
    ceres::Solver::Summary s3;
    ceres::Solve(o3, &problem, &s3);
    std::printf("%s\n", s3.BriefReport().c_str());
    bool moved = cb.seen.size() >= 3 && cb.seen[1] != cb.seen[0] && cb.seen[2] != cb.seen[1];
    out.push_back((double)s3.termination_type); out.push_back((double)s3.iterations.size()); out.push_back((double)cb.calls);
    out.push_back(moved ? 1.0 : 0.0); out.push_back(s3.IsSolutionUsable() ? 1.0 : 0.0);
  }
  FILE* f = std::fopen(argv[2], "wb");
  std::fwrite(out.data(), 8, out.size(), f);
  std::fclose(f);
  return 0;
}
