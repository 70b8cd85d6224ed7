"""CPU: synthetic generator properties, and the multi-rank sharding of the structure build (world_size 2 over gloo)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generator_is_deterministic_and_well_formed(ob):
    a = ob.synth.make_config("C1obj", seed=3)
    b = ob.synth.make_config("C1obj", seed=3)
    c = ob.synth.make_config("C1obj", seed=4)
    assert np.array_equal(a.reproj["px"], b.reproj["px"]) and np.array_equal(a.poses, b.poses)
    assert not np.array_equal(a.poses, c.poses)
    n = a.counts()
    assert n["poses"] == 50 and n["points"] == 2000 and n["objects"] == 20 and n["reproj"] > 10000 and n["bbox"] > 100
    rp = a.reproj
    order = np.lexsort((rp["point"], rp["cam"], rp["pose"]))
    assert np.array_equal(order, np.arange(len(order)))          # canonical pose-major order
    assert np.bincount(rp["point"], minlength=n["points"])[np.unique(rp["point"])].min() >= 5
    assert a.const_pose[:5].all() and not a.const_pose[5:].any()
    for cov in a.relpose["cov"]:
        assert np.all(np.diag(cov) >= 1e-6 - 1e-18)              # kMinStdDev = 1e-3 (relative_pose_factor_utils.h)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import obvi_b200 as ob
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = ob.synth.make_graph(K=60, P=1500, O=12, seed=9, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4)
    p = ob.problem_from_graph(g, device=-1)
    mine = p.debug_partition(rank, world)
    alone = p.debug_partition(0, 1)
    out = [None] * world
    dist.all_gather_object(out, mine)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put((out, alone))


def test_sharding_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    out, alone = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # observations / e-blocks are partitioned, pose-only factors live on rank 0, the reduced system is replicated
    assert sum(o["n_obs"] for o in out) == alone["n_obs"] and min(o["n_obs"] for o in out) > 0.3 * alone["n_obs"]
    assert sum(o["n_bbox"] for o in out) == alone["n_bbox"]
    assert sum(o["points_here"] for o in out) == alone["points_here"]
    assert sum(o["objects_here"] for o in out) == alone["objects_here"]
    assert out[0]["n_rel"] == alone["n_rel"] and out[1]["n_rel"] == 0
    assert sum(o["n_unary"] for o in out) == alone["n_unary"]
    for o in out:
        for k in ("nf", "n_upper", "structure_checksum", "num_parameters_reduced", "num_residual_blocks_reduced"):
            assert o[k] == alone[k], k
