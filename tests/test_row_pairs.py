"""CPU: the row-pair work lists of the point elimination (problem.hpp: row_pair_entries, item chunking) cover exactly the
(keyframe a <= keyframe b) pairs of every regular point -- counted by brute force from the graph -- for tracks with gaps,
constant poses, stereo pairs, tracks longer than the row span, constant points, and for every rank of a sharded build."""
import numpy as np
import pytest

ROW_SPAN = 20      # problem.hpp kRowSpan


def build(ob, n_pose, tracks, const_poses=(), const_points=()):
    """tracks: list of lists of (pose index, camera) observations; returns (Problem, brute-force stats)."""
    p = ob.Problem(-1)
    poses = np.zeros((n_pose, 6)); poses[:, 0] = np.arange(n_pose)
    pts = np.ones((len(tracks), 3))
    p.add_parameter_array(poses); p.add_parameter_array(pts)
    cams = [p.add_camera((400, 400, 320, 240), np.eye(3), np.array([0.1 * c, 0.0, 0.0])) for c in range(2)]
    for k in const_poses:
        p.set_parameter_block_constant(poses[k])
    for e in const_points:
        p.set_parameter_block_constant(pts[e])
    for e, tr in enumerate(tracks):
        for k, c in tr:
            p.add_reprojection(poses[k], pts[e], cams[c], (300.0 + k, 200.0 + e), 1.5, 1.0)
    # f index of every variable pose that carries at least one observation, in pose order
    used = sorted({k for tr in tracks for k, _ in tr})
    f_of = {}
    for k in used:
        if k not in const_poses:
            f_of[k] = len(f_of)
    return p, f_of


def brute(tracks, f_of, const_points=(), owned=None):
    products = slots = regular = fallback = 0
    for e, tr in enumerate(tracks):
        if owned is not None and e not in owned:
            continue
        if not tr:
            continue
        fs = sorted({f_of[k] for k, _ in tr if k in f_of})
        if e in const_points or (fs and fs[-1] - fs[0] >= ROW_SPAN):
            fallback += 1
            continue
        regular += 1
        if fs:
            slots += fs[-1] - fs[0] + 1
            products += sum(fs[-1] - a + 1 for a in fs)          # row a pairs with every dense column a .. last
    return dict(products=products, dense_slots=slots, regular_points=regular, fallback_points=fallback)


def check(ob, n_pose, tracks, const_poses=(), const_points=()):
    p, f_of = build(ob, n_pose, tracks, const_poses, const_points)
    got = p.debug_row_products(0, 1)
    want = brute(tracks, f_of, const_points)
    for k, v in want.items():
        assert got[k] == v, (k, got, want)
    assert got["entries_in_items"] == got["entries"] and got["longest_item"] <= 128
    return got


def test_small_cases(ob):
    # one stereo track over consecutive poses, starting on an even and on an odd row
    check(ob, 12, [[(k, c) for k in range(0, 7) for c in (0, 1)]])
    check(ob, 12, [[(k, c) for k in range(1, 8) for c in (0, 1)]])
    # gaps inside a track (gap poses own no row but keep a dense slot), single-pose track, two-pose track
    check(ob, 30, [[(2, 0), (3, 0), (7, 1), (8, 0), (15, 0)], [(4, 0)], [(5, 0), (6, 1)]])
    # span of exactly ROW_SPAN - 1 f indices stays regular, ROW_SPAN goes to the generic kernels
    check(ob, 40, [[(0, 0), (ROW_SPAN - 1, 0)], [(1, 0), (1 + ROW_SPAN, 0)], [(3, 0), (4, 0), (5, 0)]])
    # constant poses (no f index: the f numbering closes up), constant points
    g = check(ob, 20, [[(k, 0) for k in range(0, 10)], [(k, 1) for k in range(3, 9)], [(k, 0) for k in range(5, 12)]],
              const_poses=(0, 1, 6), const_points=(2,))
    assert g["fallback_points"] == 1
    # a point seen from constant poses only: regular, no slot, no product
    g = check(ob, 10, [[(0, 0), (1, 0)], [(2, 0), (3, 0)]], const_poses=(0, 1))
    assert g["regular_points"] == 2


def test_random_tracks(ob):
    rng = np.random.default_rng(7)
    for trial in range(6):
        n_pose = int(rng.integers(5, 70))
        tracks = []
        for e in range(int(rng.integers(1, 400))):
            k0 = int(rng.integers(0, n_pose)); L = int(rng.integers(1, 26))
            ks = [k for k in range(k0, min(n_pose, k0 + L)) if rng.random() < 0.8] or [k0]
            tracks.append([(k, c) for k in ks for c in (0, 1) if c == 0 or rng.random() < 0.7])
        cp = tuple(int(k) for k in rng.choice(n_pose, size=int(rng.integers(0, 4)), replace=False))
        cq = tuple(int(e) for e in rng.choice(len(tracks), size=min(len(tracks), int(rng.integers(0, 3))), replace=False))
        check(ob, n_pose, tracks, cp, cq)


def test_sharded_builds_partition_the_products(ob):
    g = ob.synth.make_config("C1")
    p = ob.problem_from_graph(g, device=-1)
    whole = p.debug_row_products(0, 1)
    for world in (2, 3):
        parts = [p.debug_row_products(r, world) for r in range(world)]
        for k in ("products", "dense_slots", "regular_points", "entries"):
            assert sum(q[k] for q in parts) >= whole[k] if k == "entries" else sum(q[k] for q in parts) == whole[k], (k, world)
    # small problems get short items (a warp walks its item serially), large ones the full 128 entries
    assert whole["longest_item"] <= 16 or whole["entries"] > 16 * 4096
