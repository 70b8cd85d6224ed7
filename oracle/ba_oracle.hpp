// TEST INFRASTRUCTURE ONLY -- C++ oracle (b): a CPU restatement of the ObVi-SLAM bundle-adjustment
// hot path with Ceres semantics.  PARITY UNPINNED: the reference's solver arithmetic lives in
// Ceres + SuiteSparse (un-vendored, absent from this image) and the reference has no golden
// vectors for this path, so this file is pinned only against the
// independent NumPy oracle (oracle/py_oracle.py) -- never call it "Ceres".
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the library built from this file.  Nothing under obvi-slam_b200/ links or includes it.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t K, P, O, C;
  double* poses;    // K x 6 (t, axis-angle), in/out
  double* points;   // P x 3, in/out
  double* objects;  // O x 7 (x y z yaw dx dy dz), in/out
  const uint8_t* const_pose;
  const uint8_t* const_point;
  const uint8_t* const_obj;
  const double* cam_intr;  // C x 4 (fx fy cx cy)
  const double* cam_R;     // C x 9 row-major, camera orientation in the robot frame
  const double* cam_t;     // C x 3
  int64_t n_reproj;
  const int32_t *rp_pose, *rp_point, *rp_cam;
  const double* rp_px;     // n x 2
  const double* rp_sigma;  // n
  double rp_huber;
  int64_t n_bbox;
  const int32_t *bb_obj, *bb_pose, *bb_cam;
  const double* bb_corners;  // n x 4 (xmin xmax ymin ymax)
  const double* bb_cov;      // n x 16
  double bb_huber, bb_invalid;
  int64_t n_shape;
  const int32_t* sh_obj;
  const double* sh_mean;  // n x 3
  const double* sh_cov;   // n x 9
  double sh_huber;
  int64_t n_ltm;
  const int32_t* lt_obj;
  const double* lt_mean;  // n x 7
  const double* lt_cov;   // n x 49
  double lt_huber;
  int64_t n_rel;
  const int32_t *rl_p1, *rl_p2;
  const double* rl_t;    // n x 3
  const double* rl_R;    // n x 9 row-major measured rotation
  const double* rl_cov;  // n x 36
  double rl_huber;
} oracle_graph;

typedef struct {
  int32_t max_num_iterations;
  int32_t use_nonmonotonic_steps;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_trust_region_radius, max_trust_region_radius;
  int32_t num_threads;  // <= 0: all
} oracle_options;

typedef struct {
  int32_t termination;  // 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE
  int32_t num_iterations;  // = Summary::iterations.size()
  int32_t lm_steps;        // trust-region step attempts (linear solves)
  int32_t num_parameters_reduced;
  int32_t num_threads_used;
  double initial_cost, final_cost, fixed_cost;
  double total_time, linear_solver_time, jacobian_time, residual_time;
} oracle_summary;

// per-iteration log row: iteration, cost, cost_change, step_norm, successful, radius, gradient_max_norm
#define ORACLE_LOG_COLS 7

int oracle_solve(oracle_graph* g, const oracle_options* opt, oracle_summary* out, double* log, int32_t log_rows);

// Raw or loss-corrected residuals/Jacobians per block, Ceres layout (row-major per block).
// Any output pointer may be null.  Sizes: r_reproj n*2, Jp n*12, Jl n*6; r_bbox n*4, J_obj n*28,
// J_pose n*24; r_shape n*3, J n*21; r_ltm n*7, J n*49; r_rel n*6, J1 n*36, J2 n*36.
int oracle_evaluate(const oracle_graph* g, int apply_loss, double* cost, double* r_reproj, double* jp_reproj,
                    double* jl_reproj, double* r_bbox, double* jo_bbox, double* jp_bbox, double* r_shape,
                    double* j_shape, double* r_ltm, double* j_ltm, double* r_rel, double* j1_rel, double* j2_rel);

// (Sigma^-1)^(1/2), principal root, n <= 7, row-major in/out.
void oracle_sqrt_information(const double* cov, int n, double* out);

#ifdef __cplusplus
}
#endif
