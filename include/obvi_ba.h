/*
 * obvi_ba.h -- C ABI of the B200-native nonlinear-least-squares backend for ObVi-SLAM's joint
 * keyframe-pose + 3D-point + object-ellipsoid bundle adjustment.
 *
 * This is the drop-in boundary: the entry points below are what a binding of the reference's
 * Ceres-facing path would call.  Each one cites the reference interface it replaces (paths relative
 * to the reference tree).  Plain pointers and sizes only; no C++ / torch types cross this ABI.
 * All arithmetic runs on the GPU (sm_100a); there is no CPU fallback -- every compute entry point
 * fails with OBVI_ERR_CUDA when no device is usable.
 *
 * Threading: thread-compatible, not thread-safe -- one owner thread per problem handle, exactly like
 * the reference's use of ceres::Problem.
 *
 * Memory ownership: parameter blocks stay owned by the caller (the reference's pose-graph nodes,
 * include/refactoring/optimization/low_level_feature_pose_graph.h:315-321).  The host pointer IS the
 * block identity, as with Ceres.  obvi_solve reads initial values from those pointers and writes the
 * minimum-cost iterate back to them.
 */
#ifndef OBVI_BA_H_
#define OBVI_BA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct obvi_problem obvi_problem;
typedef uint64_t obvi_factor_id; /* plays the role of ceres::ResidualBlockId */

enum obvi_status {
  OBVI_OK = 0,
  OBVI_ERR_INVALID_ARGUMENT = 1,
  OBVI_ERR_CUDA = 2,
  OBVI_ERR_NOT_FOUND = 3,
  OBVI_ERR_NUMERIC = 4, /* e.g. NaN information matrix (relative_pose_factor.cpp:14-18 exits there) */
  OBVI_ERR_COMM = 5
};

/* Factor types: the reference's kReprojectionErrorFactorTypeId .. (low_level_feature_pose_graph.h:18-23,
 * object_pose_graph.h:18-20). */
enum obvi_factor_type {
  OBVI_FACTOR_REPROJECTION = 0,
  OBVI_FACTOR_BBOX = 2,
  OBVI_FACTOR_SHAPE_PRIOR = 3,
  OBVI_FACTOR_LTM_PRIOR = 4,
  OBVI_FACTOR_REL_POSE = 5,
  OBVI_FACTOR_PARAM_PRIOR = 6
};

/* ---- problem lifetime (ceres::Problem problem; offline_problem_runner.h:113) --------------------
 * cuda_device = -1 creates a host-only handle: the problem can be assembled and inspected
 * (obvi_debug_partition), every compute call on it fails with OBVI_ERR_CUDA. */
int obvi_problem_create(int cuda_device, obvi_problem** out);
void obvi_problem_destroy(obvi_problem* p);
/* Last error text of this handle (or of creation when p == NULL). Never NULL. */
const char* obvi_last_error(const obvi_problem* p);

/* ---- parameter blocks (Problem::AddParameterBlock / RemoveParameterBlock /
 *      SetParameterBlockConstant / SetParameterBlockVariable / IsParameterBlockConstant;
 *      object_pose_graph_optimizer.h:417-422,440-472,474-613).  size: 6 = pose (t, axis-angle),
 *      3 = point, 7 = ellipsoid (x y z yaw dx dy dz).  Blocks are also added implicitly by the
 *      obvi_factor_add_* calls, as Ceres does. */
int obvi_param_add(obvi_problem* p, double* host_block, int size);
/* `count` consecutive blocks of `size` doubles starting at host_base (fast path for array-backed graphs). */
int obvi_param_add_array(obvi_problem* p, double* host_base, int size, int64_t count);
int obvi_param_remove(obvi_problem* p, double* host_block);
int obvi_param_set_constant(obvi_problem* p, double* host_block, int is_constant);
int obvi_param_is_constant(const obvi_problem* p, const double* host_block, int* is_constant);

/* ---- cameras: intrinsics (fx fy cx cy) + extrinsics = camera pose in the robot frame (R row-major, t).
 *      The factories below take them per call in the reference (CameraIntrinsicsMat, CameraExtrinsics);
 *      here they are registered once and referenced by id. */
int obvi_camera_add(obvi_problem* p, const double intrinsics_fx_fy_cx_cy[4], const double extrinsics_R[9],
                    const double extrinsics_t[3], int* camera_id);

/* ---- factors (residual_creator.h:20-436 -> Problem::AddResidualBlock(cost, new HuberLoss(a), ...)).
 *      huber <= 0 means "no loss function" (ParameterPrior is added that way,
 *      long_term_object_map_extraction.cpp:817). */
/* ReprojectionCostFunctor::create(K, extrinsics, pixel, sigma) (reprojection_cost_functor.h:135-144);
 * parameter order (pose, point) (residual_creator.h:263-264). */
int obvi_factor_add_reproj(obvi_problem* p, double* pose, double* point, int camera_id, const double pixel[2],
                           double reprojection_error_std_dev, double huber, obvi_factor_id* id);
int obvi_factor_add_reproj_batch(obvi_problem* p, int64_t n, double* const* poses, double* const* points,
                                 const int32_t* camera_ids, const double* pixels /* n x 2 */,
                                 const double* std_devs /* n */, double huber, obvi_factor_id* ids /* n or NULL */);
/* BoundingBoxFactor::createBoundingBoxFactor(invalid_err, corners (xmin xmax ymin ymax), K, extrinsics, cov4)
 * (bounding_box_factor.h:153-164); parameter order (ellipsoid, pose) (residual_creator.h:114-115). */
int obvi_factor_add_bbox(obvi_problem* p, double* ellipsoid, double* pose, int camera_id, const double corners[4],
                         const double cov4x4[16], double invalid_ellipse_error, double huber, obvi_factor_id* id);
int obvi_factor_add_bbox_batch(obvi_problem* p, int64_t n, double* const* ellipsoids, double* const* poses,
                               const int32_t* camera_ids, const double* corners /* n x 4 */,
                               const double* covs /* n x 16 */, double invalid_ellipse_error, double huber,
                               obvi_factor_id* ids);
/* ShapePriorFactor::createShapeDimPrior(mean, cov3) (shape_prior_factor.h:76-80). */
int obvi_factor_add_shape_prior(obvi_problem* p, double* ellipsoid, const double mean[3], const double cov3x3[9],
                                double huber, obvi_factor_id* id);
/* IndependentObjectMapFactor::createIndependentObjectMapFactor(mean, cov7)
 * (independent_object_map_factor.h:35-40). */
int obvi_factor_add_ltm_prior(obvi_problem* p, double* ellipsoid, const double mean[7], const double cov7x7[49],
                              double huber, obvi_factor_id* id);
/* RelativePoseFactor::createRelativePoseFactor(Pose3D measured, cov6) (relative_pose_factor.h:74-76);
 * measured rotation given as a row-major matrix. */
int obvi_factor_add_rel_pose(obvi_problem* p, double* pose_before, double* pose_after, const double measured_t[3],
                             const double measured_R[9], const double cov6x6[36], double huber, obvi_factor_id* id);
/* ParameterPrior::createParameterPrior<N>(idx, mean, std_dev) (parameter_prior.h:36-45). */
int obvi_factor_add_param_prior(obvi_problem* p, double* block, int param_idx, double mean, double std_dev,
                                double huber, obvi_factor_id* id);
/* Problem::RemoveResidualBlock (object_pose_graph_optimizer.h:1105-1155). */
int obvi_factor_remove(obvi_problem* p, obvi_factor_id id);
/* The same for n blocks in one call (the excluded-factor list between the two phases of a window holds ~10 % of its
 * residual blocks, offline_problem_runner.h:752-833); stops at the first unknown id, earlier ones stay removed. */
int obvi_factor_remove_batch(obvi_problem* p, const obvi_factor_id* ids, int64_t n);
int64_t obvi_num_factors(const obvi_problem* p);
/* How often the flat device layout was rebuilt from the host container.  Removing reprojection / bounding-box blocks of
 * an already solved problem (the outlier exclusion between the two phases, offline_problem_runner.h:752-833) is done in
 * place -- the block's record is flagged and evaluates to zero -- and does not count. */
int64_t obvi_num_structure_builds(const obvi_problem* p);
/* Problem::GetResidualBlocks: live blocks in the order obvi_evaluate concatenates them (order of addition). */
int obvi_residual_blocks(const obvi_problem* p, obvi_factor_id* ids, int32_t* types, int32_t* sizes, int64_t capacity,
                         int64_t* n);

/* ---- solve (ObjectPoseGraphOptimizer::solveOptimization, object_pose_graph_optimizer.h:634-707) -- */
typedef struct {
  /* the fields the reference sets (object_pose_graph_optimizer.h:651-672; optimization_solver_params.h:10-37) */
  int32_t max_num_iterations;
  int32_t use_nonmonotonic_steps;
  double function_tolerance;
  double gradient_tolerance;
  double parameter_tolerance;
  double initial_trust_region_radius;
  double max_trust_region_radius;
  /* Ceres defaults the reference leaves untouched */
  double min_trust_region_radius;            /* 1e-32 */
  double min_relative_decrease;              /* 1e-3 */
  double min_lm_diagonal;                    /* 1e-6 */
  double max_lm_diagonal;                    /* 1e32 */
  int32_t max_consecutive_nonmonotonic_steps; /* 5 */
  int32_t max_num_consecutive_invalid_steps;  /* 5 */
  /* reduced-camera-system solver (block-Jacobi preconditioned CG on the Schur complement) */
  int32_t pcg_max_iterations;   /* 2000 */
  double pcg_relative_tolerance; /* |r| <= tol * |b|; 1e-12 keeps the LM trajectory on the direct-solve one */
  /* ceres::IterationCallback + Solver::Options::update_state_every_iteration (object_pose_graph_optimizer.h:651-659): called on
   * the calling thread after every iteration summary, in order.  Return 0 to continue (SOLVER_CONTINUE), 1 to abort
   * (SOLVER_ABORT -> OBVI_USER_FAILURE), 2 to stop successfully (SOLVER_TERMINATE_SUCCESSFULLY -> OBVI_USER_SUCCESS).  With
   * update_state_every_iteration the caller's parameter blocks hold the current iterate when the callback runs (one extra
   * device->host copy per iteration; single-rank problems only).  NULL: no callback, the loop never leaves the device queue. */
  int32_t (*iteration_callback)(void* user, const void* iteration_summary /* obvi_iteration_summary */);
  void* iteration_callback_user;
  int32_t update_state_every_iteration;
} obvi_solver_options;

/* Ceres defaults + the struct defaults of optimization_solver_params.h:17-23 (radius 1e4 / 1e16). */
void obvi_solver_options_init(obvi_solver_options* o);

/* ceres::IterationSummary subset consumed by include/debugging/optimization_logger.h:63-74. */
typedef struct {
  int32_t iteration;
  int32_t step_is_valid;
  int32_t step_is_successful;
  int32_t linear_solver_iterations;
  double cost; /* includes fixed_cost, as Ceres logs it */
  double cost_change;
  double gradient_max_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
} obvi_iteration_summary;

/* ceres::TerminationType values */
enum obvi_termination { OBVI_CONVERGENCE = 0, OBVI_NO_CONVERGENCE = 1, OBVI_FAILURE = 2, OBVI_USER_SUCCESS = 3, OBVI_USER_FAILURE = 4 };

/* ceres::Solver::Summary subset consumed by optimization_logger.h:192-203 and solveOptimization. */
typedef struct {
  int32_t termination_type;
  int32_t num_iterations;        /* = summary.iterations.size() (includes iteration 0) */
  int32_t num_lm_steps;          /* trust-region step attempts = linear solves performed */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t num_parameter_blocks_reduced;
  int32_t num_parameters_reduced;
  int32_t num_residual_blocks_reduced;
  int32_t num_residuals_reduced;
  int32_t is_solution_usable;    /* Summary::IsSolutionUsable() */
  double initial_cost;
  double final_cost;
  double fixed_cost;
  double total_time_in_seconds;           /* wall, whole call incl. host<->device copies */
  double preprocessor_time_in_seconds;    /* structure (re)build + upload, 0 when cached */
  double linear_solver_time_in_seconds;   /* device time: Schur build + PCG + back-substitution */
  double jacobian_evaluation_time_in_seconds; /* device time */
  double residual_evaluation_time_in_seconds; /* device time */
  double minimizer_device_time_in_seconds;    /* CUDA-event time of the whole LM loop */
  int64_t pcg_iterations_total;
  int64_t kernel_launches;       /* kernels launched by this call */
} obvi_summary;

/* iterations: optional array receiving up to `capacity` iteration summaries. */
int obvi_solve(obvi_problem* p, const obvi_solver_options* options, obvi_summary* summary,
               obvi_iteration_summary* iterations, int32_t capacity);

/* ---- Problem::Evaluate (object_pose_graph_optimizer.h:679-693: apply_loss_function = false, all
 *      residual blocks) -> concatenated residuals in obvi_residual_blocks order. */
int obvi_evaluate(obvi_problem* p, int apply_loss_function, double* cost, double* residuals, int64_t capacity,
                  int64_t* num_residuals);
/* Per-type residuals and Jacobians in order of addition, Ceres layout (row-major per block):
 * reprojection r 2, J0 (pose) 2x6, J1 (point) 2x3; bbox r 4, J0 (ellipsoid) 4x7, J1 (pose) 4x6;
 * shape r 3, J0 3x7; ltm r 7, J0 7x7; rel-pose r 6, J0 6x6, J1 6x6; param prior r 1, J0 1xN padded to 7.
 * Any output may be NULL. */
int obvi_evaluate_factor_type(obvi_problem* p, int factor_type, int apply_loss_function, double* residuals,
                              double* jacobian0, double* jacobian1);

/* Problem::Evaluate with gradient / Jacobian output (long_term_object_map_extraction.cpp:251-252,591-598;
 * EvaluateOptions{apply_loss_function, residual_blocks, parameter_blocks}).  Rows: the residual blocks `ids` in that order
 * (NULL: all live blocks in order of addition); columns: the parameter blocks `blocks` in that order, constant ones left out, as
 * Ceres does (NULL: every variable block -- poses, points, ellipsoids).  Call once with crs_* = NULL to size the arrays
 * (crs_rows: num_rows + 1 entries, crs_cols / crs_values: nnz), then again to fill them.  gradient (optional, num_cols) = J^T r. */
int obvi_evaluate_jacobian(obvi_problem* p, int apply_loss_function, const obvi_factor_id* ids, int64_t n_ids,
                           double* const* blocks, int64_t n_blocks, int64_t* num_rows, int64_t* num_cols, int64_t* nnz,
                           int32_t* crs_rows, int32_t* crs_cols, double* crs_values, double* gradient);

/* ---- two-phase outlier rejection (offline_problem_runner.h:689-801): the ids of the
 *      floor(n_distinct * fraction) blocks of `factor_type` with the largest raw squared residual norm,
 *      ties collapsed as the reference's std::map<double, id, std::greater> does. */
int obvi_topk_outliers(obvi_problem* p, int factor_type, double fraction, obvi_factor_id* ids, int64_t capacity,
                       int64_t* n);

/* ---- marginal covariances of ellipsoid blocks: ceres::Covariance::Compute + GetCovarianceBlock as the long-term-map
 *      extraction uses them (src/refactoring/long_term_map/long_term_object_map_extraction.cpp:362-440; block lists in
 *      include/refactoring/long_term_map/long_term_object_map_extraction.h:269-284 (pairs) and :459-467 (diagonal)).
 *      out[i] = 7x7 row-major block [obj_a[i], obj_b[i]] of (J^T J)^-1 at the current values, loss functions applied,
 *      constant blocks left out (their blocks are zero).  Computed from the same Schur elimination as the solve, without
 *      damping; OBVI_ERR_NUMERIC when J is rank deficient (no gauge fix), as Ceres' SPARSE_QR path reports failure. */
int obvi_object_covariances(obvi_problem* p, int64_t n_pairs, double* const* obj_a, double* const* obj_b, double* out);

/* ---- multi-GPU: one process per GPU; e-blocks (points / objects) are sharded over ranks, the reduced
 *      camera system is all-reduced over NCCL each LM iteration.  unique_id is a 128-byte ncclUniqueId
 *      produced by obvi_comm_unique_id on rank 0 and distributed by the caller. */
int obvi_comm_unique_id(void* unique_id_128_bytes);
int obvi_comm_init(obvi_problem* p, const void* unique_id_128_bytes, int rank, int world_size);
/* Single-process variant: joins `world_size` handles of THIS process (handle i = rank i; same device or peer-accessible
 * devices) through host barriers + a plain reduction kernel instead of NCCL.  Each handle must then be driven by its own host
 * thread, all making the same sequence of obvi_solve / obvi_evaluate calls.  It exists so that the sharded code path can be
 * run and checked on a single-GPU box (NCCL refuses two ranks on one device); the multi-process path is obvi_comm_init. */
int obvi_comm_init_local(obvi_problem** handles, int world_size);
/* Share the communicator (and rank / world size) of `src` with another problem handle on the same device: the reference solves
 * hundreds of windows per session (offline_problem_runner.h:100-270); the communicator is created once. */
int obvi_comm_attach(obvi_problem* p, const obvi_problem* src);

/* ---- measurement hook (bench.py): times `reps` back-to-back launches of the reprojection Jacobian-evaluation
 *      kernel at the current host values with CUDA events on the solver's stream (after 3 warm-up launches) and
 *      reports the average seconds per launch and the algorithmic bytes one launch moves (SURVEY.md section 8d:
 *      192 B per observation + the unique parameter blocks it reads). */
int obvi_profile_jacobian(obvi_problem* p, int reps, double* seconds_per_launch, int64_t* algorithmic_bytes,
                          int64_t* num_observations);

/* Host-only inspection of how the structure build shards the graph for (rank, world_size); stats has 12 entries
 * (see solver.cu).  Used by the CPU-side multi-process tests. */
int obvi_debug_partition(obvi_problem* p, int rank, int world_size, int64_t* stats);
/* Host-only: the row-pair work lists of the point elimination built for (rank, world_size); stats has 8 entries: entries,
 * work items, 6x6 products the entries select (must equal the number of (keyframe a <= keyframe b) pairs over all regular
 * points), dense slots, regular points, points left to the generic kernels, entries covered by items, longest item. */
int obvi_debug_row_products(obvi_problem* p, int rank, int world_size, int64_t* stats);
/* Host-only: FNV-1a hash over every array of the structure built for (rank, world_size) -- the exact bytes the solver would
 * upload.  Lets a change of the (host) structure build be checked bit for bit without a GPU. */
int obvi_debug_structure_hash(obvi_problem* p, int rank, int world_size, uint64_t* hash);

/* Library / build information, e.g. "obvi_ba 0.1 sm_100a". */
const char* obvi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OBVI_BA_H_ */
