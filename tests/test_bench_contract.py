"""bench.py contract, CPU side: the reference arm (`--impl reference`, the Ceres-semantics CPU restatement) prints ONE JSON line
with the keys the driver reads, on rank 0 only.  (The CUDA arm needs a GPU; its line is produced by the same code path and is
recorded under profiles/.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1obj", "--steps", "3", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "LM it/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 3 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["unit"] == d["unit"]


def test_reference_arm_uses_every_host_core_under_torchrun():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm must not inherit it (round-1 SCALE ratios were against one thread)."""
    r = _run(dict(OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"))
    assert r.returncode == 0, r.stderr[-1500:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert 1 < d["cpu_baseline"]["cores"] <= len(os.sched_getaffinity(0))


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""
