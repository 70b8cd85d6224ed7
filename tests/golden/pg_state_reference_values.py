"""The hand-made pose-graph state of the reference's own I/O round-trip test, as data (values only; they are non-physical on
purpose -- ids do not line up, covariances are not symmetric, rotation axes are not unit vectors):
test/file_io/cv_file_storage/object_and_reprojection_feature_pose_graph_file_storage_io_tests.cc
  :12-20 class priors, :22-23 id range, :25-30 ellipsoids, :31-40 per-object bookkeeping, :42-45 LTM ids,
  :47-63 bounding-box factors, :65-75 shape priors, :77-113 factor index sets, :117-128 extrinsics / intrinsics,
  :131-145 ranges and robot poses, :147-195 index sets, :197-222 relative-pose factors, :223-227 reprojection factors,
  :229-241 feature bookkeeping and positions.
Factor type ids (low_level_feature_pose_graph.h:18-23, object_pose_graph.h:18-20)."""
import numpy as np

REPROJ, PAIRWISE, BBOX, SHAPE, LTM, RELPOSE = 0, 1, 2, 3, 4, 5
A = np.array


def state():
    obj = dict(
        mean_and_cov_by_semantic_class={"chair": (A([1.2, 94.3, 92.3]), A([[1.0, 2.1, 3.2], [4.3, 5.4, 6.5], [7.6, 8.7, 9.8]])),
                                        "trashcan": (A([3.2, -3.2, 18.3]), A([[1.9, 2.0, 3.1], [4.2, 5.3, 6.4], [7.5, 8.6, 9.7]]))},
        min_object_id=93, max_object_id=19038,
        ellipsoid_estimates={14: A([84.3, 913.3, 8.4, 19.3, 9.4, 58.2, 3.1]), 94: A([9.4, -184.4, 4.2, 18.3, -10.3, 4.2, 0.3])},
        semantic_class_for_object={324: "abc", 183: "def"},
        last_observed_frame_by_object={493: 139, 129: 492}, first_observed_frame_by_object={1848: 10, 19348: 193},
        min_object_observation_factor=13, max_object_observation_factor=93, min_obj_specific_factor=31, max_obj_specific_factor=193,
        long_term_map_object_ids={13, 493, 472, 846},
        object_observation_factors={
            32: dict(frame_id=94, camera_id=23, object_id=43, bounding_box_corners=A([1.2, 2.3, 3.4, 1.4]),
                     bounding_box_corners_covariance=A([[1, 2, 3, 4], [11, 12, 13, 14], [21, 22, 23, 24], [31, 32, 33, 34]], float),
                     detection_confidence=13.4),
            94: dict(frame_id=92, camera_id=91, object_id=42, bounding_box_corners=A([94.2, 42.4, 0.1, 92.1]),
                     bounding_box_corners_covariance=A([[0.1, 0.2, 0.3, 0.4], [1.1, 1.2, 1.3, 1.4], [2.1, 2.2, 2.3, 2.4], [3.1, 3.2, 3.3, 3.4]]),
                     detection_confidence=94.1)},
        shape_dim_prior_factors={
            90: dict(object_id=42, mean_shape_dim=A([4.2, 0.3, 13.3]), shape_dim_cov=A([[3.2, 45.2, 0.1], [34.1, 3.1, 0.4], [9.3, 2.5, 13.4]])),
            13: dict(object_id=135, mean_shape_dim=A([9.4, 13.4, 9.3]), shape_dim_cov=A([[0.32, 4.52, 0.01], [3.41, 0.31, 0.04], [0.93, 0.25, 1.34]]))},
        observation_factors_by_frame={42: {(REPROJ, 23)}, 91: {(SHAPE, 40), (PAIRWISE, 13)}, 194: {(LTM, 99), (BBOX, 138), (RELPOSE, 924)}},
        observation_factors_by_object={84: {(REPROJ, 3), (SHAPE, 45), (BBOX, 914)}, 76: {(LTM, 342)}, 95: {(PAIRWISE, 94842), (RELPOSE, 1345)}},
        object_only_factors_by_object={24: {(RELPOSE, 84), (SHAPE, 4567), (REPROJ, 678), (LTM, 34)}, 62: {(RELPOSE, 7892)}},
    )
    cov1 = A([[1.2, 4, 3.5, 10.4, -0.3, -20.3], [1.25, 4.5, 3.0, 11.4, -0.8, -21.3], [1.24, 4.4, 3.4, 12.4, -0.7, -22.3],
              [1.23, 4.3, 3.3, 13.4, -0.6, -23.3], [1.22, 4.2, 3.2, 14.4, -0.5, -24.3], [1.21, 4.1, 3.1, 15.4, -0.4, -25.3]])
    cov2 = A([[1.2, 2.3, 3.4, 4.5, 5.6, 6.7], [11.2, 12.3, 13.4, 14.5, 15.6, 16.7], [1.21, 2.31, 3.41, 4.51, 5.61, 6.71],
              [21.2, 22.3, 23.4, 24.5, 25.6, 26.7], [1.22, 2.32, 3.42, 4.52, 5.62, 6.72], [31.2, 32.3, 33.4, 34.5, 35.6, 36.7]])
    low = dict(
        camera_extrinsics_by_camera={1: dict(transl=A([-0.3, 4.2, 2.3]), angle=4.3, axis=A([-0.3, 12.3, -9.0])),
                                     2: dict(transl=A([-1.3, 7.2, -2.3]), angle=413.0, axis=A([-0.13, 142.3, -9.1]))},
        camera_intrinsics_by_camera={1: A([[3.2, 89.3, 0.2], [1.4, 3.4, 9.3], [0.5, 0.2, 1.3]]), 2: A([[13.2, 19.3, 1.2], [2.4, 6.4, 8.3], [1.5, 9.2, 1.5]])},
        visual_factor_type=REPROJ, min_frame_id=0, max_frame_id=500, max_feature_factor_id=9825256, max_pose_factor_id=135,
        robot_poses={1: A([1.2, 2.3, 3.4, 4.5, 5.6, 6.7]), 2: A([1.3, 2.4, 3.5, 4.6, 5.7, 6.8]), 5: A([1.4, 2.5, 3.6, 4.7, 5.8, 6.9])},
        pose_factors_by_frame={10: {(REPROJ, 12), (PAIRWISE, 72)}, 510: {(RELPOSE, 973)}, 190: {(REPROJ, 10384), (PAIRWISE, 384), (BBOX, 104)}},
        visual_feature_factors_by_frame={284: [(REPROJ, 13), (BBOX, 420)], 953: [(LTM, 134)], 344: [(LTM, 42), (BBOX, 3), (RELPOSE, 948)]},
        visual_factors_by_feature={24: {(RELPOSE, 21)}, 94: {(BBOX, 124), (LTM, 13)}, 301: {(PAIRWISE, 139), (SHAPE, 938), (REPROJ, 492)}},
        pose_factors={123: dict(frame_id_1=1, frame_id_2=2, measured_pose_deviation=dict(transl=A([4.2, 0.4, -0.3]), angle=-np.pi, axis=A([0.4, -19.3, 48.2])),
                                pose_deviation_cov=cov1),
                      94: dict(frame_id_1=3, frame_id_2=4, measured_pose_deviation=dict(transl=A([4.6, 0.2, -9.4]), angle=-np.pi / 3, axis=A([-9.3, 34.2, -0.2])),
                               pose_deviation_cov=cov2)},
        factors={32: dict(frame_id=1, feature_id=2, camera_id=3, feature_pos=A([1.2, 3.4]), reprojection_error_std_dev=4.2),
                 832: dict(frame_id=4, feature_id=3, camera_id=49, feature_pos=A([-38.4, 39.4]), reprojection_error_std_dev=1.3)},
        last_observed_frame_by_feature={4: 1, 38: 183, 188: 973}, first_observed_frame_by_feature={5: 2, 39: 184, 189: 974},
    )
    return dict(low=low, min_feature_id=10, max_feature_id=50,
                feature_positions={5: A([1.2, 3.4, 5.6]), 6: A([2.3, 4.5, 6.7]), 7: A([-0.35, -483.3, 9.2])}, obj=obj)
