"""CPU: the C-ABI library loads, exports every symbol include/obvi_ba.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "obvi_ba.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(obvi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(ob):
    lib = ob.lib()
    names = declared_symbols()
    assert len(names) >= 29
    for n in names:
        assert hasattr(lib, n), f"libobvi_ba.so does not export {n}"
    assert set(ob.EXPORTED_SYMBOLS) <= set(names)
    assert b"sm_100a" in lib.obvi_version()


def test_no_cpu_fallback(ob):
    """Without a CUDA device the product refuses to compute (it never routes through the oracle)."""
    import shutil
    if shutil.which("nvidia-smi") and os.system("nvidia-smi -L > /dev/null 2>&1") == 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ob.ObviError):
        ob.Problem(0)
    p = ob.Problem(-1)  # host-only handle: assembly works, compute fails loudly
    poses = np.zeros((2, 6)); pts = np.ones((1, 3))
    p.add_parameter_array(poses); p.add_parameter_array(pts)
    cam = p.add_camera((400, 400, 320, 240), np.eye(3), np.zeros(3))
    p.add_reprojection(poses[0], pts[0], cam, (300.0, 200.0), 1.5, 1.0)
    assert p.num_residual_blocks() == 1
    with pytest.raises(ob.ObviError, match="no CUDA device|no CPU fallback"):
        p.solve(max_num_iterations=1)
    with pytest.raises(ob.ObviError):
        p.evaluate()


def test_problem_bookkeeping_host_only(ob):
    """Add / remove / constant bookkeeping mirrors ceres::Problem (no compute involved)."""
    p = ob.Problem(-1)
    poses = np.zeros((3, 6)); pts = np.ones((4, 3)); objs = np.ones((1, 7))
    for a in (poses, pts, objs):
        p.add_parameter_array(a)
    cam = p.add_camera((400, 400, 320, 240), np.eye(3), np.zeros(3))
    assert p.add_camera((400, 400, 320, 240), np.eye(3), np.zeros(3)) == cam  # deduplicated
    ids = [p.add_reprojection(poses[i % 3], pts[i % 4], cam, (10.0 * i, 5.0), 1.5, 1.0) for i in range(6)]
    b = p.add_bounding_box(objs[0], poses[1], cam, (1, 2, 3, 4), np.eye(4) * 900, 1000.0, 0.5)
    s = p.add_shape_prior(objs[0], (1, 1, 1), np.eye(3), 10.0)
    r = p.add_relative_pose(poses[0], poses[1], np.zeros(3), np.eye(3), np.eye(6), 1.0)
    pr = p.add_parameter_prior(objs[0], 3, 0.1, 0.5)
    assert p.num_residual_blocks() == 10
    got, types, sizes = p.residual_blocks()
    assert list(got) == ids + [b, s, r, pr] and list(sizes) == [2] * 6 + [4, 3, 6, 1]
    assert list(types) == [0] * 6 + [2, 3, 5, 6]
    p.remove_residual_block(ids[2])
    with pytest.raises(ob.ObviError):
        p.remove_residual_block(ids[2])
    assert p.num_residual_blocks() == 9 and ids[2] not in p.residual_blocks()[0]
    # the batch call removes like a loop of single calls and stops at the first unknown id
    extra = [p.add_reprojection(poses[0], pts[1], cam, (1.0 + k, 2.0), 1.5, 1.0) for k in range(3)]
    assert p.num_residual_blocks() == 12
    p.remove_residual_blocks(extra[:2])
    assert p.num_residual_blocks() == 10
    with pytest.raises(ob.ObviError):
        p.remove_residual_blocks([extra[2], extra[0]])
    assert p.num_residual_blocks() == 9 and extra[2] not in p.residual_blocks()[0]
    p.remove_residual_blocks([])
    assert not p.is_parameter_block_constant(poses[0])
    p.set_parameter_block_constant(poses[0])
    assert p.is_parameter_block_constant(poses[0])
    p.set_parameter_block_variable(poses[0])
    assert not p.is_parameter_block_constant(poses[0])
    p.remove_parameter_block(pts[0])      # removes the residual blocks that use it, like Ceres
    assert p.num_residual_blocks() == 7
    with pytest.raises(ob.ObviError):
        p.add_relative_pose(poses[0], poses[1], np.zeros(3), np.eye(3), np.full((6, 6), np.nan), 1.0)
    st = p.debug_partition(0, 1)
    assert st["nf"] == 3 and st["n_obs"] == 3 and st["n_bbox"] == 1 and st["n_rel"] == 1 and st["n_unary"] == 2


def test_recycled_block_address_and_array_overlap(ob):
    """ADVICE r1 (problem.hpp): a removed block's address may come back with another size (Ceres accepts the sequence); an array
    may not be registered over a block that is already registered singly anywhere inside its range."""
    p = ob.Problem(-1)
    buf = np.zeros(8)
    pose, cam_pts = np.zeros(6), np.ones((2, 3))
    p.add_parameter_block(buf[:6])                 # a 6-block at this address ...
    p.remove_parameter_block(buf[:6])
    p.add_parameter_block(buf[:3])                 # ... recycled as a 3-block
    cam = p.add_camera((400, 400, 320, 240), np.eye(3), np.zeros(3))
    p.add_parameter_block(pose)
    p.add_reprojection(pose, buf[:3], cam, (300.0, 200.0), 1.5, 1.0)
    assert p.debug_partition(0, 1)["n_obs"] == 1
    with pytest.raises(ob.ObviError):              # a LIVE block keeps its size
        p.add_parameter_block(buf[:7])
    # one row registered singly, referenced by a factor, then the whole array: refused (the factor would keep the old id)
    q = ob.Problem(-1)
    pts = np.ones((5, 3)); pose2 = np.zeros(6)
    cam = q.add_camera((400, 400, 320, 240), np.eye(3), np.zeros(3))
    q.add_parameter_block(pts[2]); q.add_parameter_block(pose2)
    q.add_reprojection(pose2, pts[2], cam, (300.0, 200.0), 1.5, 1.0)
    with pytest.raises(ob.ObviError, match="overlaps"):
        q.add_parameter_array(pts)
    q.set_parameter_block_constant(pts[2])
    assert q.is_parameter_block_constant(pts[2])
    other = np.ones((4, 3))
    q.add_parameter_array(other)                   # disjoint arrays are fine; so is re-adding rows of a registered array
    q.add_parameter_block(other[1])
    with pytest.raises(ob.ObviError, match="overlaps"):
        q.add_parameter_array(other[1:3])
