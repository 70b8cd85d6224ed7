"""Reader for the reference's simulated visual-SLAM sequences (data/vslam_set1..7, data/vslam_superset1/<density>/<noise>):
a directory of per-frame text files + features/features.txt + calibration/camera_matrix.txt.  Format, as specified by the
data sets' own README (data/vslam_set4/README.md: "Frame Coordinates", "Label Format", "Data File", "Calibration"):

    NNNNNN.txt                     line 1  frame id
                                   line 2  map-frame pose of the robot  x y z qx qy qz qw
                                   rest    feature_id  u  v            (pixels; ids consistent across frames)
    features/features.txt          feature_id  X Y Z                  (3-D landmark, map frame)
    calibration/camera_matrix.txt  fx fy cx cy

The poses are those of a body frame with x forward; the single camera looks along it with the usual optical axes (z forward, x
right, y down) and no offset -- with that convention the noise-free set (vslam_set7) reprojects its landmarks onto its keypoints
to 1e-6 px, which tests/test_vslam_dataset.py checks.  The result is the same plain-NumPy factor graph the synthetic generator
produces (synth.FactorGraph): reprojection factors only, poses as (translation, angle-axis) like the reference's RawPose3d
(vslam_basic_types_refactor.h:18-80).
"""
import glob
import os

import numpy as np

from . import synth

R_BODY_FROM_OPTICAL = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])   # camera axes in the body frame (columns)


def _quat_to_angle_axis(q):
    """(qx, qy, qz, qw) -> angle * axis with angle in [0, pi]; exact at pi, where the rotation-matrix log map degenerates
    (vslam_set7 has a frame turned by exactly pi)."""
    q = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    if q[3] < 0.0:
        q = -q
    n = np.linalg.norm(q[0:3])
    if n < 1e-300:
        return np.zeros(3)
    return q[0:3] / n * (2.0 * np.arctan2(n, q[3]))


def read_vslam_dataset(directory, sigma_px=1.5, huber=1.0, min_observations=2):
    """-> (graph, frame_ids, feature_ids).  Landmarks seen in fewer than `min_observations` frames are dropped together with
    their keypoints (the reference's `min_visual_feature_parallax_*` / minimum-observation gates do the same job upstream)."""
    intr = tuple(float(v) for v in open(os.path.join(directory, "calibration", "camera_matrix.txt")).read().split()[:4])
    feats = {}
    for line in open(os.path.join(directory, "features", "features.txt")):
        f = line.split()
        if len(f) >= 4:
            feats[int(f[0])] = np.array([float(f[1]), float(f[2]), float(f[3])])
    frames = []
    for path in sorted(glob.glob(os.path.join(directory, "[0-9]*.txt"))):
        lines = [l for l in open(path).read().splitlines() if l.strip()]
        fid = int(lines[0].split()[0])
        pose = np.array(lines[1].split(), dtype=np.float64)
        obs = [(int(l.split()[0]), float(l.split()[1]), float(l.split()[2])) for l in lines[2:]]
        frames.append((fid, pose, obs))
    frames.sort(key=lambda f: f[0])
    count = {}
    for _, _, obs in frames:
        for i, _, _ in obs:
            if i in feats:
                count[i] = count.get(i, 0) + 1
    keep = sorted(i for i, c in count.items() if c >= min_observations)
    index = {i: n for n, i in enumerate(keep)}
    g = synth.FactorGraph()
    g.poses = np.zeros((len(frames), 6))
    for k, (_, pose, _) in enumerate(frames):
        g.poses[k, 0:3] = pose[0:3]
        g.poses[k, 3:6] = _quat_to_angle_axis(pose[3:7])
    g.points = np.stack([feats[i] for i in keep]) if keep else np.zeros((0, 3))
    g.objects = np.zeros((0, 7))
    g.cams = [dict(intr=intr, R=R_BODY_FROM_OPTICAL.copy(), t=np.zeros(3))]
    op, oq, px = [], [], []
    for k, (_, _, obs) in enumerate(frames):
        for i, u, v in obs:
            if i in index:
                op.append(k); oq.append(index[i]); px.append((u, v))
    op, oq = np.array(op, np.int64), np.array(oq, np.int64)
    order = np.lexsort((oq, op))
    g.reproj = dict(pose=op[order], point=oq[order], cam=np.zeros(len(op), np.int64), px=np.ascontiguousarray(np.array(px).reshape(-1, 2)[order]),
                    sigma=np.full(len(op), float(sigma_px)), huber=float(huber))
    g.bbox = dict(obj=np.zeros(0, np.int64), pose=np.zeros(0, np.int64), cam=np.zeros(0, np.int64), corners=np.zeros((0, 4)),
                  cov=np.zeros((0, 4, 4)), huber=0.5, invalid_err=1000.0)
    g.shape = dict(obj=np.zeros(0, np.int64), mean=np.zeros((0, 3)), cov=np.zeros((0, 3, 3)), huber=10.0)
    g.ltm = dict(obj=np.zeros(0, np.int64), mean=np.zeros((0, 7)), cov=np.zeros((0, 7, 7)), huber=1.0)
    g.relpose = dict(p1=np.zeros(0, np.int64), p2=np.zeros(0, np.int64), t=np.zeros((0, 3)), Rm=np.zeros((0, 3, 3)), cov=np.zeros((0, 6, 6)), huber=1.0)
    g.const_pose = np.zeros(len(frames), dtype=bool)
    g.const_point = np.zeros(len(keep), dtype=bool)
    g.const_obj = np.zeros(0, dtype=bool)
    return g, [f[0] for f in frames], keep
