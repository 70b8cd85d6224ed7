// Drop-in replacement of include/refactoring/factors/bounding_box_factor.h (createBoundingBoxFactor signature as in
// bounding_box_factor.h:149-178).  Registers a bbox factor with the CUDA backend (csrc/factors.cuh:
// bbox_residual_jacobian, analytic 4x7 / 4x6 Jacobians; invalid-ellipse branch -> constant residual, zero Jacobian).
#ifndef UT_VSLAM_BOUNDING_BOX_FACTOR_H
#define UT_VSLAM_BOUNDING_BOX_FACTOR_H

#include <ceres/autodiff_cost_function.h>
#include <refactoring/types/vslam_basic_types_refactor.h>
#include <refactoring/types/vslam_obj_opt_types_refactor.h>

#include <optional>

#include "obvi_factor_common.h"

namespace vslam_types_refactor {

class BoundingBoxFactor {
 public:
  BoundingBoxFactor(const double& invalid_ellipse_error, const CameraIntrinsicsMat<double>& intrinsics,
                    const CameraExtrinsics<double>& extrinsics, const BbCorners<double>& corner_pixel_locations,
                    const Covariance<double, 4>& corner_detections_covariance, const std::optional<ObjectId>& obj_id = std::nullopt,
                    const std::optional<FrameId>& frame_id = std::nullopt, const std::optional<CameraId>& camera_id = std::nullopt,
                    const bool& debug = false)
      : invalid_ellipse_error_(invalid_ellipse_error), camera_(obvi_shim::makeCamera(intrinsics, extrinsics)),
        obj_id_(obj_id), frame_id_(frame_id), camera_id_(camera_id), debug_(debug) {
    for (int i = 0; i < 4; i++) corners_[i] = corner_pixel_locations(i);  // (xmin, xmax, ymin, ymax)
    obvi_shim::copySquare<4>(corner_detections_covariance, cov_);
  }

  // parameter order (ellipsoid, pose), residual_creator.h:114-115
  int obviAdd(obvi_problem* p, double* const* blocks, double huber, obvi_factor_id* id) const {
    int cam = -1;
    const int rc = obvi_shim::registerCamera(p, camera_, &cam);
    if (rc != OBVI_OK) return rc;
    return obvi_factor_add_bbox(p, blocks[0], blocks[1], cam, corners_, cov_, invalid_ellipse_error_, huber, id);
  }

  static ceres::AutoDiffCostFunction<BoundingBoxFactor, 4, kEllipsoidParamterizationSize, 6>* createBoundingBoxFactor(
      const double& invalid_ellipse_error, const vslam_types_refactor::BbCorners<double>& object_detection,
      const vslam_types_refactor::CameraIntrinsicsMat<double>& camera_intrinsics,
      const vslam_types_refactor::CameraExtrinsics<double>& camera_extrinsics,
      const vslam_types_refactor::Covariance<double, 4>& bounding_box_covariance, const std::optional<ObjectId>& obj_id,
      const std::optional<FrameId>& frame_id, const std::optional<CameraId>& cam_id, const bool& debug = false) {
    BoundingBoxFactor* factor = new BoundingBoxFactor(invalid_ellipse_error, camera_intrinsics, camera_extrinsics, object_detection,
                                                      bounding_box_covariance, obj_id, frame_id, cam_id, debug);
    return new ceres::AutoDiffCostFunction<BoundingBoxFactor, 4, kEllipsoidParamterizationSize, 6>(factor);
  }

 private:
  double invalid_ellipse_error_;
  obvi_shim::CameraData camera_;
  double corners_[4];
  double cov_[16];
  std::optional<ObjectId> obj_id_;
  std::optional<FrameId> frame_id_;
  std::optional<CameraId> camera_id_;
  bool debug_;
};
}  // namespace vslam_types_refactor
#endif  // UT_VSLAM_BOUNDING_BOX_FACTOR_H
