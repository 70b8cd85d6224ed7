"""Reader / writer of the reference's pose-graph-state JSON.

Format = what `outputPoseGraphStateToFile` / `readPoseGraphStateFromFile` produce and consume through cv::FileStorage in
JSON mode (include/file_io/cv_file_storage/object_and_reprojection_feature_pose_graph_file_storage_io.h:1021-1046), i.e.
the `Serializable*` classes of that header plus the primitives of file_storage_io_utils.h and
vslam_basic_types_file_storage_io.h / vslam_obj_types_file_storage_io.h:

  top level                 {"pose_graph": {"reprojection_low_level_feature_pose_graph_state": {...}, "obj_only_pose_graph_state_": {...}}}
  SerializableMap           [{"k": key, "v": value}, ...]                      (file_storage_io_utils.h:42-71)
  SerializablePair          {"f": first, "s": second}                          (:123-145)
  SerializableVector        [{"i": index, "v": value}, ...]                    (:186-219)
  Serializable*Set          [entry, ...]                                       (:243-315)
  SerializableUint64        the decimal STRING of the id (frame / feature / camera / object / factor ids)   (:441-457)
  SerializableEigenMat      {"Rows": r, "Cols": c, "Data": [row-major values]}  (vslam_basic_types_file_storage_io.h:22-66)
  SerializablePose3D        {"transl": mat 3x1, "rot": {"angle": a, "axis": mat 3x1}}   (:91-170)
  raw pose / ellipsoid      6x1 / 7x1 matrices (t, axis-angle) / (x y z yaw dx dy dz)
  factor types              integers: 0 reprojection, 2 bounding box, 3 shape prior, 5 relative pose
                            (low_level_feature_pose_graph.h:18-23, object_pose_graph.h:18-20)

OpenCV is not available in this environment, so the dialect is transcribed from those headers, not validated against a
file written by cv::FileStorage; the reader is tolerant (ids may be strings or numbers, derived index maps are ignored and
rebuilt from the factors), the writer emits every key the reference's reader asks for.  The long-term-map priors are not part
of this file (the reference keeps them in the long-term-map file); `long_term_map_object_ids` is carried through.
"""
from __future__ import annotations

import json

import numpy as np

from . import synth

TOP = "pose_graph"
K_LOW = "reprojection_low_level_feature_pose_graph_state"
K_OBJ = "obj_only_pose_graph_state_"
T_REPROJ, T_BBOX, T_SHAPE, T_RELPOSE = 0, 2, 3, 5


# ----------------------------------------------------------------------------- primitives
def _mat(a):
    a = np.asarray(a, dtype=np.float64)
    assert a.ndim == 2
    return {"Rows": int(a.shape[0]), "Cols": int(a.shape[1]), "Data": [float(v) for v in a.ravel()]}


def _vec(a):
    a = np.asarray(a, dtype=np.float64).ravel()
    return {"Rows": int(a.size), "Cols": 1, "Data": [float(v) for v in a]}


def _unmat(n):
    return np.asarray(n["Data"], dtype=np.float64).reshape(int(n["Rows"]), int(n["Cols"]))


def _uid(v):
    return str(int(v))


def _id(n):
    return int(n)


def _map(pairs):
    return [{"k": k, "v": v} for k, v in pairs]


def _unmap(n):
    return [(e["k"], e["v"]) for e in (n or [])]


def _pose3d(t, R):
    rv = synth.mat_to_rotvec(np.asarray(R)[None])[0]
    ang = float(np.linalg.norm(rv))
    axis = rv / ang if ang > 0 else np.array([1.0, 0.0, 0.0])
    return {"transl": _vec(t), "rot": {"angle": ang, "axis": _vec(axis)}}


def _unpose3d(n):
    t = _unmat(n["transl"]).ravel()
    rv = float(n["rot"]["angle"]) * _unmat(n["rot"]["axis"]).ravel()
    return t, synth.rotvec_to_mat(rv[None])[0]


def _fset(entries):
    return [{"f": int(t), "s": _uid(i)} for t, i in entries]


# ----------------------------------------------------------------------------- writer
def write_pose_graph_state(path, g, ids=None, semantic_classes=None, class_priors=None):
    """Write a FactorGraph as the reference's pose-graph state.  `ids` (optional, as returned by the reader) keeps the
    original frame / feature / object / camera / factor ids; otherwise indices are used."""
    ids = ids or {}
    frame = np.asarray(ids.get("frame", np.arange(len(g.poses))), dtype=np.int64)
    feat = np.asarray(ids.get("feature", np.arange(len(g.points))), dtype=np.int64)
    obj = np.asarray(ids.get("object", np.arange(len(g.objects))), dtype=np.int64)
    cam = np.asarray(ids.get("camera", np.arange(len(g.cams))), dtype=np.int64)
    rp, bb, sh, rl = g.reproj, g.bbox, g.shape, g.relpose
    n_rp, n_bb, n_sh, n_rl = len(rp["pose"]), len(bb["obj"]), len(sh["obj"]), len(rl["p1"])
    f_rp = np.asarray(ids.get("reproj_factor", np.arange(n_rp)), dtype=np.int64)
    f_rl = np.asarray(ids.get("relpose_factor", np.arange(n_rl)), dtype=np.int64)
    f_bb = np.asarray(ids.get("bbox_factor", np.arange(n_bb)), dtype=np.int64)
    f_sh = np.asarray(ids.get("shape_factor", np.arange(n_sh)), dtype=np.int64)

    def grouped(keys, types, fids):
        out = {}
        for k, t, f in zip(keys, types, fids):
            out.setdefault(int(k), []).append((int(t), int(f)))
        return out

    low = {
        "camera_extrinsics_by_camera": _map((_uid(cam[c]), _pose3d(g.cams[c]["t"], g.cams[c]["R"])) for c in range(len(g.cams))),
        "camera_intrinsics_by_camera": _map((_uid(cam[c]), _mat(np.array([[g.cams[c]["intr"][0], 0.0, g.cams[c]["intr"][2]],
                                                                           [0.0, g.cams[c]["intr"][1], g.cams[c]["intr"][3]],
                                                                           [0.0, 0.0, 1.0]]))) for c in range(len(g.cams))),
        "visual_factor_type": T_REPROJ,
        "min_frame_id": _uid(frame.min() if len(frame) else 0),
        "max_frame_id": _uid(frame.max() if len(frame) else 0),
        "max_feature_factor_id": _uid(f_rp.max() if n_rp else 0),
        "max_pose_factor_id": _uid(f_rl.max() if n_rl else 0),
        "robot_poses": _map((_uid(frame[k]), _vec(g.poses[k])) for k in range(len(g.poses))),
        "pose_factors_by_frame": _map((_uid(k), _fset(v)) for k, v in grouped(
            np.concatenate([frame[rl["p1"]], frame[rl["p2"]]]) if n_rl else [], [T_RELPOSE] * (2 * n_rl), np.concatenate([f_rl, f_rl]) if n_rl else []).items()),
        "visual_feature_factors_by_frame": _map((_uid(k), [{"i": i, "v": e} for i, e in enumerate(_fset(v))]) for k, v in grouped(
            frame[rp["pose"]] if n_rp else [], [T_REPROJ] * n_rp, f_rp).items()),
        "visual_factors_by_feature": _map((_uid(k), _fset(v)) for k, v in grouped(feat[rp["point"]] if n_rp else [], [T_REPROJ] * n_rp, f_rp).items()),
        "pose_factors": _map((_uid(f_rl[i]), {"frame_id_1": _uid(frame[rl["p1"][i]]), "frame_id_2": _uid(frame[rl["p2"][i]]),
                                             "measured_pose_deviation": _pose3d(rl["t"][i], rl["Rm"][i]), "pose_deviation_cov": _mat(rl["cov"][i])})
                             for i in range(n_rl)),
        "factors": _map((_uid(f_rp[i]), {"frame_id": _uid(frame[rp["pose"][i]]), "feature_id": _uid(feat[rp["point"][i]]),
                                        "camera_id": _uid(cam[rp["cam"][i]]), "feature_pos": _vec(rp["px"][i]),
                                        "reprojection_error_std_dev": float(rp["sigma"][i])}) for i in range(n_rp)),
    }
    first, last = {}, {}
    for p_, k in zip(rp["point"], rp["pose"]):
        a, b = int(feat[p_]), int(frame[k])
        first[a] = min(first.get(a, b), b); last[a] = max(last.get(a, b), b)
    low["last_observed_frame_by_feature"] = _map((_uid(k), _uid(v)) for k, v in last.items())
    low["first_observed_frame_by_feature"] = _map((_uid(k), _uid(v)) for k, v in first.items())
    used_pts = sorted(set(int(p_) for p_ in rp["point"]))
    reproj_state = {"low_level_pg_state": low, "min_feature_id": _uid(feat.min() if len(feat) else 0),
                    "max_feature_id": _uid(feat.max() if len(feat) else 0),
                    "feature_positions": _map((_uid(feat[i]), _vec(g.points[i])) for i in (used_pts if ids.get("only_observed_features") else range(len(g.points))))}

    ofirst, olast = {}, {}
    for o, k in zip(bb["obj"], bb["pose"]):
        a, b = int(obj[o]), int(frame[k])
        ofirst[a] = min(ofirst.get(a, b), b); olast[a] = max(olast.get(a, b), b)
    semantic_classes = semantic_classes or {}
    obj_state = {
        "mean_and_cov_by_semantic_class": _map((str(c), {"f": _vec(m), "s": _mat(cv)}) for c, (m, cv) in (class_priors or {}).items()),
        "min_object_id": _uid(obj.min() if len(obj) else 0), "max_object_id": _uid(obj.max() if len(obj) else 0),
        "ellipsoid_estimates": _map((_uid(obj[o]), _vec(g.objects[o])) for o in range(len(g.objects))),
        "semantic_class_for_object": _map((_uid(obj[o]), str(semantic_classes.get(int(obj[o]), ""))) for o in range(len(g.objects))),
        "last_observed_frame_by_object": _map((_uid(k), _uid(v)) for k, v in olast.items()),
        "first_observed_frame_by_object": _map((_uid(k), _uid(v)) for k, v in ofirst.items()),
        "min_object_observation_factor": _uid(f_bb.min() if n_bb else 0), "max_object_observation_factor": _uid(f_bb.max() if n_bb else 0),
        "min_obj_specific_factor": _uid(f_sh.min() if n_sh else 0), "max_obj_specific_factor": _uid(f_sh.max() if n_sh else 0),
        "long_term_map_object_ids": [_uid(o) for o in ids.get("ltm_objects", [])],
        "object_observation_factors": _map((_uid(f_bb[i]), {"frame_id": _uid(frame[bb["pose"][i]]), "camera_id": _uid(cam[bb["cam"][i]]),
                                                           "object_id": _uid(obj[bb["obj"][i]]), "bounding_box_corners": _vec(bb["corners"][i]),
                                                           "bounding_box_corners_covariance": _mat(bb["cov"][i]),
                                                           "detection_confidence": float(ids.get("bbox_confidence", np.ones(n_bb))[i])}) for i in range(n_bb)),
        "shape_dim_prior_factors": _map((_uid(f_sh[i]), {"object_id": _uid(obj[sh["obj"][i]]), "mean_shape_dim": _vec(sh["mean"][i]),
                                                        "shape_dim_cov": _mat(sh["cov"][i])}) for i in range(n_sh)),
        "observation_factors_by_frame": _map((_uid(k), _fset(v)) for k, v in grouped(frame[bb["pose"]] if n_bb else [], [T_BBOX] * n_bb, f_bb).items()),
        "observation_factors_by_object": _map((_uid(k), _fset(v)) for k, v in grouped(obj[bb["obj"]] if n_bb else [], [T_BBOX] * n_bb, f_bb).items()),
        "object_only_factors_by_object": _map((_uid(k), _fset(v)) for k, v in grouped(obj[sh["obj"]] if n_sh else [], [T_SHAPE] * n_sh, f_sh).items()),
    }
    with open(path, "w") as f:
        json.dump({TOP: {K_LOW: reproj_state, K_OBJ: obj_state}}, f)


# ----------------------------------------------------------------------------- reader
def read_pose_graph_state(path, huber=None):
    """Read a pose-graph state into a FactorGraph (+ the id tables needed to write it back).  `huber` overrides the loss
    parameters (the file does not hold them: they live in the configuration; defaults = config/base7a_2_fallback.json)."""
    hub = dict(reproj=1.0, bbox=0.5, shape=10.0, relpose=1.0, ltm=1.0, invalid_err=1000.0)
    hub.update(huber or {})
    with open(path) as f:
        root = json.load(f)[TOP]
    rs, os_ = root[K_LOW], root[K_OBJ]
    low = rs["low_level_pg_state"]
    poses = sorted((_id(k), _unmat(v).ravel()) for k, v in _unmap(low["robot_poses"]))
    feats = sorted((_id(k), _unmat(v).ravel()) for k, v in _unmap(rs["feature_positions"]))
    objs = sorted((_id(k), _unmat(v).ravel()) for k, v in _unmap(os_["ellipsoid_estimates"]))
    intr = {_id(k): _unmat(v) for k, v in _unmap(low["camera_intrinsics_by_camera"])}
    extr = {_id(k): _unpose3d(v) for k, v in _unmap(low["camera_extrinsics_by_camera"])}
    cam_ids = sorted(intr)
    frame_of = {k: i for i, (k, _) in enumerate(poses)}
    feat_of = {k: i for i, (k, _) in enumerate(feats)}
    obj_of = {k: i for i, (k, _) in enumerate(objs)}
    cam_of = {k: i for i, k in enumerate(cam_ids)}
    g = synth.FactorGraph()
    g.poses = np.ascontiguousarray([p for _, p in poses], dtype=np.float64).reshape(-1, 6)
    g.points = np.ascontiguousarray([p for _, p in feats], dtype=np.float64).reshape(-1, 3)
    g.objects = np.ascontiguousarray([p for _, p in objs], dtype=np.float64).reshape(-1, 7)
    g.cams = [dict(intr=(intr[c][0, 0], intr[c][1, 1], intr[c][0, 2], intr[c][1, 2]), R=extr[c][1], t=extr[c][0]) for c in cam_ids]
    fac = sorted((_id(k), v) for k, v in _unmap(low["factors"]))
    g.reproj = dict(pose=np.array([frame_of[_id(v["frame_id"])] for _, v in fac], np.int64), point=np.array([feat_of[_id(v["feature_id"])] for _, v in fac], np.int64),
                    cam=np.array([cam_of[_id(v["camera_id"])] for _, v in fac], np.int64),
                    px=np.array([_unmat(v["feature_pos"]).ravel() for _, v in fac], np.float64).reshape(-1, 2),
                    sigma=np.array([float(v["reprojection_error_std_dev"]) for _, v in fac], np.float64), huber=hub["reproj"])
    pf = sorted((_id(k), v) for k, v in _unmap(low["pose_factors"]))
    tr = [_unpose3d(v["measured_pose_deviation"]) for _, v in pf]
    g.relpose = dict(p1=np.array([frame_of[_id(v["frame_id_1"])] for _, v in pf], np.int64), p2=np.array([frame_of[_id(v["frame_id_2"])] for _, v in pf], np.int64),
                     t=np.array([t for t, _ in tr], np.float64).reshape(-1, 3), Rm=np.array([R for _, R in tr], np.float64).reshape(-1, 3, 3),
                     cov=np.array([_unmat(v["pose_deviation_cov"]) for _, v in pf], np.float64).reshape(-1, 6, 6), huber=hub["relpose"])
    of = sorted((_id(k), v) for k, v in _unmap(os_["object_observation_factors"]))
    g.bbox = dict(obj=np.array([obj_of[_id(v["object_id"])] for _, v in of], np.int64), pose=np.array([frame_of[_id(v["frame_id"])] for _, v in of], np.int64),
                  cam=np.array([cam_of[_id(v["camera_id"])] for _, v in of], np.int64),
                  corners=np.array([_unmat(v["bounding_box_corners"]).ravel() for _, v in of], np.float64).reshape(-1, 4),
                  cov=np.array([_unmat(v["bounding_box_corners_covariance"]) for _, v in of], np.float64).reshape(-1, 4, 4),
                  huber=hub["bbox"], invalid_err=hub["invalid_err"])
    sf = sorted((_id(k), v) for k, v in _unmap(os_["shape_dim_prior_factors"]))
    g.shape = dict(obj=np.array([obj_of[_id(v["object_id"])] for _, v in sf], np.int64),
                   mean=np.array([_unmat(v["mean_shape_dim"]).ravel() for _, v in sf], np.float64).reshape(-1, 3),
                   cov=np.array([_unmat(v["shape_dim_cov"]) for _, v in sf], np.float64).reshape(-1, 3, 3), huber=hub["shape"])
    g.ltm = dict(obj=np.zeros(0, np.int64), mean=np.zeros((0, 7)), cov=np.zeros((0, 7, 7)), huber=hub["ltm"])
    g.const_pose = np.zeros(len(g.poses), bool); g.const_point = np.zeros(len(g.points), bool); g.const_obj = np.zeros(len(g.objects), bool)
    ids = dict(frame=np.array([k for k, _ in poses], np.int64), feature=np.array([k for k, _ in feats], np.int64),
               object=np.array([k for k, _ in objs], np.int64), camera=np.array(cam_ids, np.int64),
               reproj_factor=np.array([k for k, _ in fac], np.int64), relpose_factor=np.array([k for k, _ in pf], np.int64),
               bbox_factor=np.array([k for k, _ in of], np.int64), shape_factor=np.array([k for k, _ in sf], np.int64),
               bbox_confidence=np.array([float(v.get("detection_confidence", 1.0)) for _, v in of]),
               ltm_objects=[_id(o) for o in os_.get("long_term_map_object_ids", [])])
    extras = dict(semantic_classes={_id(k): v for k, v in _unmap(os_.get("semantic_class_for_object"))},
                  class_priors={k: (_unmat(v["f"]).ravel(), _unmat(v["s"])) for k, v in _unmap(os_.get("mean_and_cov_by_semantic_class"))})
    return g, ids, extras


# ----------------------------------------------------------------------------- state-level API
# The file as the reference's own structs hold it (ObjectAndReprojectionFeaturePoseGraphState =
# ReprojectionLowLevelFeaturePoseGraphState{LowLevelFeaturePoseGraphState} + ObjOnlyPoseGraphState), without the
# consistency a FactorGraph needs: ids are arbitrary uint64, factors may reference frames / features that have no estimate,
# a Pose3D keeps its (angle, axis) pair verbatim (Eigen::AngleAxis does not normalise).  Python mirror:
#   uint64-keyed maps -> dict[int, ...];  sets of (factor type, factor id) -> set[tuple[int, int]];  the one vector of such
#   pairs -> list[tuple];  matrices -> numpy arrays;  Pose3D -> dict(transl, angle, axis);  factor structs -> dicts with the
#   reference's field names (trailing underscore dropped).
# This is what the reference's round-trip test exercises with hand-made values
# (test/file_io/cv_file_storage/object_and_reprojection_feature_pose_graph_file_storage_io_tests.cc:9-262).
def _enc_pose3d(p):
    return {"transl": _vec(p["transl"]), "rot": {"angle": float(p["angle"]), "axis": _vec(p["axis"])}}


def _dec_pose3d(n):
    return dict(transl=_unmat(n["transl"]).ravel(), angle=float(n["rot"]["angle"]), axis=_unmat(n["rot"]["axis"]).ravel())


def _enc_idmap(m, enc):
    return _map((_uid(k), enc(v)) for k, v in m.items())


def _dec_idmap(n, dec):
    return {_id(k): dec(v) for k, v in _unmap(n)}


def _dec_fset(n):
    return {(int(e["f"]), _id(e["s"])) for e in (n or [])}


_RELPOSE_F = (("frame_id_1", _uid, _id), ("frame_id_2", _uid, _id), ("measured_pose_deviation", _enc_pose3d, _dec_pose3d),
              ("pose_deviation_cov", _mat, _unmat))
_REPROJ_F = (("frame_id", _uid, _id), ("feature_id", _uid, _id), ("camera_id", _uid, _id), ("feature_pos", _vec, lambda n: _unmat(n).ravel()),
             ("reprojection_error_std_dev", float, float))
_BBOX_F = (("frame_id", _uid, _id), ("camera_id", _uid, _id), ("object_id", _uid, _id), ("bounding_box_corners", _vec, lambda n: _unmat(n).ravel()),
           ("bounding_box_corners_covariance", _mat, _unmat), ("detection_confidence", float, float))
_SHAPE_F = (("object_id", _uid, _id), ("mean_shape_dim", _vec, lambda n: _unmat(n).ravel()), ("shape_dim_cov", _mat, _unmat))


def _enc_struct(fields):
    return lambda d: {k: enc(d[k]) for k, enc, _ in fields}


def _dec_struct(fields):
    return lambda n: {k: dec(n[k]) for k, _, dec in fields}


_vec_dec = lambda n: _unmat(n).ravel()
_fset_map = (lambda m: _enc_idmap(m, lambda s: _fset(sorted(s))), lambda n: _dec_idmap(n, _dec_fset))
_id_map = (lambda m: _enc_idmap(m, _uid), lambda n: _dec_idmap(n, _id))
# (key, encoder, decoder) in the order the reference writes them (…file_storage_io.h:270-352, 604-616, 681-760)
_LOW_FIELDS = (
    ("camera_extrinsics_by_camera", lambda m: _enc_idmap(m, _enc_pose3d), lambda n: _dec_idmap(n, _dec_pose3d)),
    ("camera_intrinsics_by_camera", lambda m: _enc_idmap(m, _mat), lambda n: _dec_idmap(n, _unmat)),
    ("visual_factor_type", int, int), ("min_frame_id", _uid, _id), ("max_frame_id", _uid, _id),
    ("max_feature_factor_id", _uid, _id), ("max_pose_factor_id", _uid, _id),
    ("robot_poses", lambda m: _enc_idmap(m, _vec), lambda n: _dec_idmap(n, _vec_dec)),
    ("pose_factors_by_frame",) + _fset_map,
    ("visual_feature_factors_by_frame", lambda m: _enc_idmap(m, lambda v: [{"i": i, "v": e} for i, e in enumerate(_fset(v))]),
     lambda n: _dec_idmap(n, lambda v: [(int(e["v"]["f"]), _id(e["v"]["s"])) for e in sorted(v, key=lambda e: int(e["i"]))])),
    ("visual_factors_by_feature",) + _fset_map,
    ("pose_factors", lambda m: _enc_idmap(m, _enc_struct(_RELPOSE_F)), lambda n: _dec_idmap(n, _dec_struct(_RELPOSE_F))),
    ("factors", lambda m: _enc_idmap(m, _enc_struct(_REPROJ_F)), lambda n: _dec_idmap(n, _dec_struct(_REPROJ_F))),
    ("last_observed_frame_by_feature",) + _id_map, ("first_observed_frame_by_feature",) + _id_map,
)
_OBJ_FIELDS = (
    ("mean_and_cov_by_semantic_class", lambda m: _map((str(c), {"f": _vec(v[0]), "s": _mat(v[1])}) for c, v in m.items()),
     lambda n: {k: (_vec_dec(v["f"]), _unmat(v["s"])) for k, v in _unmap(n)}),
    ("min_object_id", _uid, _id), ("max_object_id", _uid, _id),
    ("ellipsoid_estimates", lambda m: _enc_idmap(m, _vec), lambda n: _dec_idmap(n, _vec_dec)),
    ("semantic_class_for_object", lambda m: _enc_idmap(m, str), lambda n: _dec_idmap(n, str)),
    ("last_observed_frame_by_object",) + _id_map, ("first_observed_frame_by_object",) + _id_map,
    ("min_object_observation_factor", _uid, _id), ("max_object_observation_factor", _uid, _id),
    ("min_obj_specific_factor", _uid, _id), ("max_obj_specific_factor", _uid, _id),
    ("long_term_map_object_ids", lambda s: [_uid(o) for o in sorted(s)], lambda n: {_id(o) for o in (n or [])}),
    ("object_observation_factors", lambda m: _enc_idmap(m, _enc_struct(_BBOX_F)), lambda n: _dec_idmap(n, _dec_struct(_BBOX_F))),
    ("shape_dim_prior_factors", lambda m: _enc_idmap(m, _enc_struct(_SHAPE_F)), lambda n: _dec_idmap(n, _dec_struct(_SHAPE_F))),
    ("observation_factors_by_frame",) + _fset_map, ("observation_factors_by_object",) + _fset_map,
    ("object_only_factors_by_object",) + _fset_map,
)


def write_state(path, state):
    """state = dict(low=..., min_feature_id=, max_feature_id=, feature_positions={id: xyz}, obj=...) -- see the comment above."""
    low = {k: enc(state["low"][k]) for k, enc, _ in _LOW_FIELDS}
    rs = {"low_level_pg_state": low, "min_feature_id": _uid(state["min_feature_id"]), "max_feature_id": _uid(state["max_feature_id"]),
          "feature_positions": _enc_idmap(state["feature_positions"], _vec)}
    obj = {k: enc(state["obj"][k]) for k, enc, _ in _OBJ_FIELDS}
    with open(path, "w") as f:
        json.dump({TOP: {K_LOW: rs, K_OBJ: obj}}, f)


def read_state(path):
    with open(path) as f:
        root = json.load(f)[TOP]
    rs = root[K_LOW]
    return dict(low={k: dec(rs["low_level_pg_state"][k]) for k, _, dec in _LOW_FIELDS}, min_feature_id=_id(rs["min_feature_id"]),
                max_feature_id=_id(rs["max_feature_id"]), feature_positions=_dec_idmap(rs["feature_positions"], _vec_dec),
                obj={k: dec(root[K_OBJ][k]) for k, _, dec in _OBJ_FIELDS})


def states_equal(a, b):
    """Field-by-field equality (exact, as the reference's operator== on these structs)."""
    if isinstance(a, dict):
        return isinstance(b, dict) and set(a) == set(b) and all(states_equal(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)) and not (a and isinstance(a[0], (int, float)) and isinstance(a, tuple)):
        return isinstance(b, (list, tuple)) and len(a) == len(b) and all(states_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.shape(a) == np.shape(b) and bool(np.array_equal(np.asarray(a), np.asarray(b)))
    return a == b
