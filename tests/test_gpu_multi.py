"""GPU, needs >= 2 devices (skipped otherwise): the sharded solve (e-blocks dealt to ranks, NCCL all-reduce of the
reduced system) must reproduce the single-GPU solve."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import obvi_b200 as ob
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
g = ob.synth.make_graph(K=60, P=3000, O=12, seed=9, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4, ltm_frac=0.3)
g1 = g.copy()
o = dict(max_num_iterations=8, function_tolerance=1e-6, initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)
p = ob.problem_from_graph(g, device=rank)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.tensor(list(ob.Problem.comm_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
p.comm_init(bytes(uid.cpu().tolist()), rank, world)
s = p.solve(**o)
out = dict(rank=rank, costs=[it["cost"] for it in s.iterations], final=s.final_cost)
if rank == 0:
    p1 = ob.problem_from_graph(g1, device=0)
    s1 = p1.solve(**o)
    out.update(single=[it["cost"] for it in s1.iterations], single_final=s1.final_cost,
               dpose=float(np.abs(g.poses - g1.poses).max()), dpoint=float(np.abs(g.points - g1.points).max()),
               dobj=float(np.abs(g.objects - g1.objects).max()))
    print("RESULT " + json.dumps(out))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_solve_matches_single_gpu(tmp_path):
    try:
        n = int(subprocess.check_output(["nvidia-smi", "-L"]).decode().count("GPU "))
    except Exception:
        n = 0
    if n < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    r = json.loads(line[len("RESULT "):])
    assert len(r["costs"]) == len(r["single"])
    for a, b in zip(r["costs"], r["single"]):
        assert abs(a - b) <= 1e-7 * abs(b)
    assert abs(r["final"] - r["single_final"]) <= 1e-7 * r["single_final"]
    assert r["dpose"] < 1e-6 and r["dpoint"] < 1e-5 and r["dobj"] < 1e-5


# ---- the same sharded code path on ONE device: `world` handles of this process, one host thread each, joined by
#      obvi_comm_init_local (host barriers + a reduction kernel in place of NCCL).  Runs on any GPU box.
def _local_sharded_solve(ob, g, world, opts, env=None):
    import threading
    graphs = [g.copy() for _ in range(world)]
    probs = [ob.problem_from_graph(gr, device=0) for gr in graphs]
    ob.Problem.comm_init_local(probs)
    out, err = [None] * world, [None] * world

    def run(r):
        try:
            out[r] = probs[r].solve(**opts)
        except Exception as e:          # a rank that throws would leave the others at a barrier: surface it
            err[r] = e
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in th), "sharded solve hung (ranks disagreed on a collective)"
    assert all(e is None for e in err), err
    return graphs, probs, out


OPTS = dict(max_num_iterations=10, function_tolerance=1e-6, initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_solve_on_one_device_matches_unsharded(ob, world):
    g = ob.synth.make_graph(K=60, P=3000, O=12, seed=9, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4, ltm_frac=0.3)
    g1 = g.copy()
    s1 = ob.problem_from_graph(g1, device=0).solve(**OPTS)
    graphs, probs, out = _local_sharded_solve(ob, g, world, OPTS)
    # every rank holds a strict subset of the observations, and the shards partition them
    n_obs = [p.debug_partition(r, world)["n_obs"] for r, p in enumerate(probs)]
    assert sum(n_obs) == len(g.reproj["pose"]) and all(0 < n < len(g.reproj["pose"]) for n in n_obs)
    for r in range(world):
        s = out[r]
        assert s.termination == s1.termination and len(s.iterations) == len(s1.iterations)
        for a, b in zip(s.iterations, s1.iterations):
            assert abs(a["cost"] - b["cost"]) <= 1e-7 * abs(b["cost"]) and a["successful"] == b["successful"]
        assert abs(s.final_cost - s1.final_cost) <= 1e-7 * s1.final_cost
        # every rank writes back the full, merged solution
        import numpy as np
        assert np.abs(graphs[r].poses - g1.poses).max() < 1e-6 and np.abs(graphs[r].points - g1.points).max() < 1e-5
        assert np.abs(graphs[r].objects - g1.objects).max() < 1e-5


@pytest.mark.parametrize("fail_rank", [0, 1])
def test_sharded_ranks_agree_on_the_preconditioner_fallback(ob, fail_rank, monkeypatch):
    """the solver status (PCG break / iterations / factorisation failure) must be rank 0's on every rank, otherwise
    ranks issue different collectives.  A failed factorisation is forced on one rank: with rank 0 failing, every rank redoes the
    step with block-Jacobi PCG; with rank 1 failing, nobody does (rank 0's solution is the one that is used).  Either way the
    solve completes and matches the unsharded one."""
    g = ob.synth.make_graph(K=40, P=1500, O=6, seed=11, objects_on=True, relpose="all", n_const_poses=1, min_obj_obs=4)
    g1 = g.copy()
    s1 = ob.problem_from_graph(g1, device=0).solve(**OPTS)
    monkeypatch.setenv("OBVI_DEBUG_BT_FAIL_RANK", str(fail_rank))
    graphs, probs, out = _local_sharded_solve(ob, g, 2, OPTS)
    for s in out:
        assert len(s.iterations) == len(s1.iterations)
        assert abs(s.final_cost - s1.final_cost) <= 1e-6 * s1.final_cost
    # rank 0's failure sends every step through the fallback: many more (block-Jacobi) PCG iterations, identically on both ranks
    assert out[0].pcg_iterations_total == out[1].pcg_iterations_total
    if fail_rank == 0:
        assert out[0].pcg_iterations_total > 3 * s1.pcg_iterations_total
