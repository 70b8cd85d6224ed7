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


def test_structure_build_is_independent_of_thread_count(tmp_path):
    """The structure build (csrc/problem.hpp) runs its passes over the factors in parallel (chunked stable counting sort,
    atomic min / or / add); every array it produces must be bit-identical to the one-thread build, for the whole problem and for
    every rank's share of a sharded one.  obvi_debug_structure_hash hashes exactly the bytes the solver would upload."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "hash.py"
    script.write_text(f"""
import sys, json
sys.path.insert(0, {root!r})
import numpy as np
import obvi_b200 as ob
out = []
g1 = ob.synth.make_graph(K=120, P=6000, O=12, seed=3, objects_on=True, relpose="starved", n_const_poses=5, min_obj_obs=4, ltm_frac=0.3, max_obj_kf=100)
g1.const_point[::7] = True; g1.const_obj[1] = True
g2 = ob.synth.make_graph(K=40, P=1500, O=6, seed=61, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4)
perm = np.random.default_rng(0).permutation(len(g2.reproj["pose"]))
for k in ("pose", "point", "cam", "px", "sigma"):
    g2.reproj[k] = np.ascontiguousarray(g2.reproj[k][perm])
for g in (g1, g2):
    p = ob.problem_from_graph(g, device=-1)
    out.append([p.debug_structure_hash(0, 1)] + [p.debug_structure_hash(r, 3) for r in range(3)])
    for fid in p.factor_ids["reproj"][::13][:200]:
        p.remove_residual_block(fid)
    out[-1].append(p.debug_structure_hash(0, 1))
print("RESULT " + json.dumps(out))
""")
    def run(threads):
        r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, OMP_NUM_THREADS=str(threads)), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        return json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    one, many = run(1), run(5)
    assert one == many
    assert len(set(one[0])) == len(one[0])          # different ranks / edits give different structures
