"""CPU: the product's analytic factor Jacobians (csrc/factors.cuh compiled for the host by tests/hostcheck) against the
dual-number autodiff of the C++ oracle, at random and adversarial points."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
_d = C.POINTER(C.c_double)


def p(a):
    return a.ctypes.data_as(_d)


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck")], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))
    lib.hc_huber.restype = C.c_double
    lib.hc_huber.argtypes = [C.c_double, C.c_double, _d]
    return lib


@pytest.fixture(scope="module")
def graph(ob):
    g = ob.synth.make_graph(30, 400, 12, seed=3, objects_on=True, relpose="all", n_const_poses=1, min_point_obs=3, min_obj_obs=4,
                            min_bbox_px=10.0, min_parallax_deg=0.0)
    g.poses[3, 3:6] = [3e-9, 0, 0]          # below the 1e-8 branch: constant identity, zero rotation Jacobian
    g.poses[4, 3:6] = [2e-8, 1e-9, 0]       # just above
    g.poses[5, 3:6] = 0.0
    g.poses[6, 3:6] = np.array([0.6, 0.5, 0.62]) / np.linalg.norm([0.6, 0.5, 0.62]) * (np.pi - 1e-3)
    g.poses[7, 3:6] = [1e-5, 2e-5, -1e-5]
    return g


def test_reprojection_jacobian(hc, oracle, graph):
    g, ev, rp = graph, oracle.evaluate(graph), graph.reproj
    assert len(rp["pose"]) > 1000
    err = np.zeros(3)
    for n in range(len(rp["pose"])):
        cam = g.cams[rp["cam"][n]]
        r, Jp, Jl = np.zeros(2), np.zeros((2, 6)), np.zeros((2, 3))
        hc.hc_reproj(p(g.poses[rp["pose"][n]]), p(g.points[rp["point"][n]]), p(rp["px"][n]), p(np.array(cam["intr"])), p(cam["R"]),
                     p(cam["t"]), C.c_double(rp["sigma"][n]), p(r), p(Jp), p(Jl))
        err = np.maximum(err, [rel_err(r, ev["r_reproj"][n]), rel_err(Jp, ev["jp_reproj"][n]), rel_err(Jl, ev["jl_reproj"][n])])
    assert err.max() < 1e-11, err
    # the translation block of the pose Jacobian is exactly -J_point: the device stores the chunk without it
    # (csrc/ba_kernels.cuh decode_chunk) and rebuilds it bit for bit
    for n in range(0, len(rp["pose"]), 7):
        cam = g.cams[rp["cam"][n]]
        r, Jp, Jl = np.zeros(2), np.zeros((2, 6)), np.zeros((2, 3))
        hc.hc_reproj(p(g.poses[rp["pose"][n]]), p(g.points[rp["point"][n]]), p(rp["px"][n]), p(np.array(cam["intr"])), p(cam["R"]),
                     p(cam["t"]), C.c_double(rp["sigma"][n]), p(r), p(Jp), p(Jl))
        assert np.array_equal(Jp[:, :3], -Jl)
    small = np.isin(rp["pose"], [3, 5])   # constant-identity branch: rotation columns are exactly zero
    assert small.any() and np.all(ev["jp_reproj"][small][:, :, 3:] == 0.0)


def test_bbox_jacobian_and_invalid_branch(hc, oracle, graph):
    g, ev, bb = graph, oracle.evaluate(graph), graph.bbox
    assert len(bb["obj"]) > 50
    err = np.zeros(3)
    for n in range(len(bb["obj"])):
        cam = g.cams[bb["cam"][n]]
        r, Jo, Jp = np.zeros(4), np.zeros((4, 7)), np.zeros((4, 6))
        hc.hc_bbox(p(g.objects[bb["obj"][n]]), p(g.poses[bb["pose"][n]]), p(bb["corners"][n]), p(bb["cov"][n]), p(np.array(cam["intr"])),
                   p(cam["R"]), p(cam["t"]), C.c_double(1000.0), p(r), p(Jo), p(Jp))
        err = np.maximum(err, [rel_err(r, ev["r_bbox"][n]), rel_err(Jo, ev["jo_bbox"][n]), rel_err(Jp, ev["jp_bbox"][n])])
    assert err.max() < 1e-10, err
    # camera inside the ellipsoid: an inner sqrt argument is <= 0 -> constant residual, zero Jacobian
    cam = g.cams[0]
    ell = np.array([0.0, 0.0, 0.0, 0.3, 50.0, 50.0, 50.0]); pose = np.array([0.1, 0.2, 0.0, 0.01, 0.02, 0.03])
    r, Jo, Jp = np.zeros(4), np.ones((4, 7)), np.ones((4, 6))
    hc.hc_bbox(p(ell), p(pose), p(np.array([100.0, 200.0, 100.0, 200.0])), p(np.eye(4) * 900.0), p(np.array(cam["intr"])), p(cam["R"]),
               p(cam["t"]), C.c_double(1000.0), p(r), p(Jo), p(Jp))
    assert np.all(r == 1000.0) and not Jo.any() and not Jp.any()


def test_relative_pose_jacobian(hc, oracle, graph):
    g, ev, rl = graph, oracle.evaluate(graph), graph.relpose
    err = np.zeros(3)
    for n in range(len(rl["p1"])):
        r, J1, J2 = np.zeros(6), np.zeros((6, 6)), np.zeros((6, 6))
        hc.hc_relpose(p(g.poses[rl["p1"][n]]), p(g.poses[rl["p2"][n]]), p(rl["t"][n]), p(rl["Rm"][n]), p(rl["cov"][n]), p(r), p(J1), p(J2))
        err = np.maximum(err, [rel_err(r, ev["r_rel"][n]), rel_err(J1, ev["j1_rel"][n]), rel_err(J2, ev["j2_rel"][n])])
    assert err.max() < 1e-10, err


def test_relative_pose_exact_identity_error(hc, oracle, ob):
    """PGO builds its measurements from the current estimates (pose_graph_plus_objects_optimizer.h:108-119): the rotation
    error is the identity up to rounding; both implementations must agree there too."""
    g = ob.synth.make_graph(6, 10, 0, seed=4, objects_on=False, relpose="all", n_const_poses=1, pose_noise=False, min_point_obs=3,
                            min_parallax_deg=0.0)
    ev, rl = oracle.evaluate(g), g.relpose
    for n in range(len(rl["p1"])):
        r, J1, J2 = np.zeros(6), np.zeros((6, 6)), np.zeros((6, 6))
        hc.hc_relpose(p(g.poses[rl["p1"][n]]), p(g.poses[rl["p2"][n]]), p(rl["t"][n]), p(rl["Rm"][n]), p(rl["cov"][n]), p(r), p(J1), p(J2))
        assert np.abs(r).max() < 1e-9
        assert np.abs(r - ev["r_rel"][n]).max() < 1e-12
        assert rel_err(J1, ev["j1_rel"][n]) < 1e-6 and rel_err(J2, ev["j2_rel"][n]) < 1e-6


def test_host_math(hc, oracle):
    rng = np.random.default_rng(1)
    for n in (3, 4, 6, 7):
        A = rng.normal(size=(n, n)); cov = np.ascontiguousarray(A @ A.T + n * np.eye(n)); out = np.zeros((n, n))
        hc.hc_sqrt_information(p(cov), n, p(out))
        assert np.allclose(out, oracle.sqrt_information(cov), rtol=1e-11, atol=1e-13)
    A = rng.normal(size=(7, 7)); M = np.ascontiguousarray(A @ A.T + 7 * np.eye(7)); inv = np.zeros((7, 7))
    assert hc.hc_spd_inverse7(p(M), p(inv)) == 1 and np.allclose(inv, np.linalg.inv(M), rtol=1e-11)
    assert hc.hc_spd_inverse3(p(np.ascontiguousarray(-np.eye(3))), p(np.zeros((3, 3)))) == 0
    sc = C.c_double(0)
    assert hc.hc_huber(1.0, 0.25, C.byref(sc)) == 0.125 and sc.value == 1.0
    assert hc.hc_huber(1.0, 9.0, C.byref(sc)) == pytest.approx(2.5) and sc.value == pytest.approx(np.sqrt(1.0 / 3.0))
