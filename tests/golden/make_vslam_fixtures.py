"""Converts simulated sequences that ship with the reference (data/vslam_set4, data/vslam_set7,
data/vslam_superset1/{low,high}_density/high_noise; format: data/vslam_set4/README.md) into small .npz fixtures, so that tests can
use the reference's OWN input data on machines where /root/reference does not exist.  The fixtures hold the files' content as read
by obvi-slam_b200/vslam_dataset_io.py (robot poses as translation + angle-axis, ground-truth landmarks seen at least twice,
keypoints), nothing computed.  Run here:  python tests/golden/make_vslam_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import obvi_b200 as ob  # noqa: E402
from make_golden import graph_arrays  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/data"
SETS = {"vslam_set7": "vslam_set7", "vslam_set4": "vslam_set4",
        "vslam_superset1_low_density_high_noise": "vslam_superset1/low_density/high_noise",
        "vslam_superset1_high_density_high_noise": "vslam_superset1/high_density/high_noise"}


def main():
    for name, rel in SETS.items():
        g, frame_ids, feature_ids = ob.vslam_dataset_io.read_vslam_dataset(os.path.join(DATA, rel))
        d = graph_arrays(g)
        d["frame_ids"] = np.array(frame_ids, np.int64)
        d["feature_ids"] = np.array(feature_ids, np.int64)
        d["source"] = np.array("data/" + rel)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, g.counts(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
