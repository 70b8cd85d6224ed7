#!/usr/bin/env python
"""bench.py -- LM iterations/s of the global bundle adjustment (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (through the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the Ceres-semantics restatement (oracle/)

A "step" is one Levenberg-Marquardt iteration (one linear solve + one candidate-cost evaluation, plus one
Jacobian evaluation when the step is accepted) on the synthetic graph S(2000, 200000, 500, seed 0) with the full
residual set (BASELINE.json configs[2]); W warm-up iterations, then EXACTLY K timed iterations from the same
initial point with all tolerances disabled.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LM iterations/sec (global BA, 2k KF / 200k pts / 500 obj)"
UNIT = "LM it/s"
WORKLOAD = dict(workload="C3: 2000 keyframes / 200k points / 500 objects, full residual set (reproj + bbox + shape-prior + rel-pose)",
                generator="obvi-slam_b200/synth.py make_config('C3', seed=0)",
                solver="LM (Ceres semantics), radius 100 / max 1e4, non-monotonic, Huber 1.0/0.5/10/1.0, tolerances disabled for timing",
                l2="working set per iteration (0.09 GB observations + 0.36 GB Jacobian chunks) exceeds the 126 MB L2: no flush needed")


JAC_KERNEL = "reproj_jac_fused_kernel (reprojection residual + Jacobian evaluation, point-major scatter, fused pose-side sums)"
TRAFFIC_FILE = "r02_jacobian_traffic.json"

_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes its version banner to stdout when
    NCCL_DEBUG is set in the environment, as it is on the GPU boxes), so file descriptor 1 is pointed at stderr for the whole run
    and the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def workload(args):
    """`config` of the JSON line; the named workload is C3 (BASELINE.json configs[2]) -- say so when another one was asked for."""
    if args.config == "C3":
        return dict(WORKLOAD)
    return dict(WORKLOAD, workload=f"{args.config} (NOT the benchmark workload; obvi-slam_b200/synth.py make_config('{args.config}'))",
                generator=f"obvi-slam_b200/synth.py make_config('{args.config}', seed={args.seed})")


def solver_opts(iters):
    return dict(max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0,
                initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.t0, self.t1 = [], None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05] or [r for _, r in self.rows]
        sm, smax, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))


def measured_peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    """Host threads this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, so the count is taken
    from the affinity mask -- capped by the cgroup CPU quota when one is set (a container that shows 24 CPUs but is allowed 6
    CPU-seconds per second runs 24 OpenMP threads far slower than 6) -- and handed to the oracle explicitly (it calls
    omp_set_num_threads, oracle/ba_oracle.cpp)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]          # cgroup v2
        if quota != "max":
            n = max(1, min(n, -(-int(quota) // int(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())   # v1
            if q > 0:
                n = max(1, min(n, -(-q // per)))
        except Exception:
            pass
    return n


_THREADS = {}


def pick_threads(g):
    """Thread count for the CPU restatement, MEASURED: one LM iteration of this graph with host_cores(), half and a quarter of
    it (and 32 / 16 on boxes with more cores than that), the fastest wins.  On the multi-GPU boxes the visible core count (24) is not what the process gets -- 24 OpenMP
    threads ran the solve 20x slower than 16 threads on the single-GPU box, with no cgroup quota to read -- so the count is
    not trusted, it is timed (a few seconds).  The timings travel in the JSON line (cpu_baseline.threads_tried)."""
    key = id(g)
    if key not in _THREADS:
        n = host_cores()
        tried = {}
        for t in sorted({n, max(1, n // 2), max(1, n // 4), min(n, 32), min(n, 16)}, reverse=True):
            tried[t] = cpu_leg(g, 1, threads=t)[2]
            if len(tried) > 1 and tried[t] > 1.5 * min(tried.values()):
                break                                   # getting slower with fewer threads: stop
        _THREADS[key] = (min(tried, key=tried.get), {str(k): round(v, 3) for k, v in tried.items()})
    return _THREADS[key]


def cpu_leg(g, iters, threads=None, tolerances=False):
    """The Ceres-semantics CPU restatement (oracle/ba_oracle.cpp) on the same graph, on all host cores; returns
    (it/s, summary dict, loop seconds, the solved copy of the graph).  tolerances=True: the config's own termination tests."""
    from oracle import oracle_lib
    gc = g.copy()
    tol = dict(function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8) if tolerances else \
        dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    r = oracle_lib.solve(gc, max_num_iterations=iters, initial_radius=100.0, max_radius=1e4, use_nonmonotonic_steps=True,
                         num_threads=threads or pick_threads(g)[0], **tol)
    loop = r["jacobian_time"] + r["linear_solver_time"] + r["residual_time"]
    return r["lm_steps"] / loop, r, loop, gc


def run_reference(args, rank):
    if rank != 0:
        return
    import obvi_b200 as ob
    g = ob.synth.make_config(args.config, seed=args.seed)
    if args.warmup > 0:
        cpu_leg(g, min(args.warmup, 1))
    t0 = time.time()
    v, r, loop, _ = cpu_leg(g, args.steps)
    cores = r["num_threads"]
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=r["lm_steps"], warmup=args.warmup,
                ms_per_step=1e3 * loop / r["lm_steps"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", config=workload(args),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{r['lm_steps']} LM iterations of the full {args.config} graph, Ceres-semantics restatement "
                                         f"(dual-number autodiff, Schur, sparse Cholesky), OpenMP {cores} threads; "
                                         f"jac {r['jacobian_time']:.2f}s lin {r['linear_solver_time']:.2f}s res {r['residual_time']:.2f}s",
                                  threads_tried=pick_threads(g)[1]),
                e2e=dict(value=r["lm_steps"] / (time.time() - t0), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="Ceres + SuiteSparse are not installable here (SURVEY.md 8c): this arm is the CPU restatement, not a Ceres binary")
    emit(line)


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import obvi_b200 as ob
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = ob.synth.make_config(args.config, seed=args.seed)
    p = ob.problem_from_graph(g, device=local_rank)
    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(ob.Problem.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        p.comm_init(bytes(uid.cpu().tolist()), rank, world)
    x0 = (g.poses.copy(), g.points.copy(), g.objects.copy())

    def reset():
        g.poses[:], g.points[:], g.objects[:] = x0

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # warm-up: W untimed iterations; this FIRST call also builds + uploads the structure, like the first Solve on a ceres::Problem
    # (what a one-shot global BA pays: reported as e2e_cold)
    barrier()
    t_first = time.time()
    s_first = p.solve(**solver_opts(max(args.warmup, 1)))
    t_first = time.time() - t_first
    reset()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.2)
    barrier()
    t0 = time.time()
    s = p.solve(**solver_opts(args.steps))  # the call synchronises the device before returning
    t1 = time.time()
    barrier()
    dev_t, wall_t = s.minimizer_device_time_in_seconds, t1 - t0
    if dist is not None:
        import torch
        t = torch.tensor([dev_t, wall_t, t_first], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_t, wall_t, t_first = t.tolist()
    clocks = None
    if sampler:
        time.sleep(0.1)
        sampler.stop()
        clocks = sampler.summary(t0, t1)
    steps = s.num_lm_steps
    gpu_cost = s.iterations[-1]["cost"] if s.iterations else float("nan")
    # parity leg (every N): the same graph solved to TERMINATION under the config's own tolerances (final-BA block of
    # config/base7a_2_fallback.json:64-87: 300 iterations, ftol 1e-6, gtol 1e-10, ptol 1e-8, radius 100 / 1e4, non-monotonic),
    # compared below with the CPU oracle run the same way: termination, iteration count, final cost, pose translations
    reset()
    sp = p.solve(max_num_iterations=300, function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8,
                 initial_trust_region_radius=100.0, max_trust_region_radius=1e4, use_nonmonotonic_steps=1)
    gpu_poses = g.poses.copy()
    roof = cpu = parity = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        reset()
        sec, nbytes, nobs = p.profile_jacobian(reps=30)     # rank 0's shard of the observations when N > 1 (no collective involved)
        roof = dict(bound="hbm", kernel=JAC_KERNEL,
                    achieved=nbytes / sec / 1e9, peak=peak, unit="GB/s", frac=nbytes / sec / 1e9 / peak, traffic=None,
                    peak_source=peak_src, algorithmic_bytes_per_launch=nbytes, observations=nobs, us_per_launch=sec * 1e6,
                    note="algorithmic bytes = SURVEY 8(d): 192 B per observation + parameter blocks once; live CUDA-event timing of 30 launches")
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
            if world == 1 and prof.get("observations") == nobs:
                roof["traffic"] = prof.get("dram_bytes_per_launch")
            roof["traffic_source"] = f"profiles/{TRAFFIC_FILE}: one ncu --set full capture of this kernel on this workload (static, not measured in this run)"
        except Exception:
            pass
        # CPU baseline = the oracle solving the same graph to termination on all host cores (a bounded sample: ~15 iterations)
        v, r, loop, gc = cpu_leg(g, 300, tolerances=True)
        cpu = dict(value=v, unit=UNIT, cores=r["num_threads"], kind="port",
                   sample=f"{r['lm_steps']} LM iterations (the whole solve to termination under the config's tolerances) of the full "
                          f"{args.config} graph, Ceres-semantics restatement (oracle/ba_oracle.cpp), {loop:.1f} s of CPU work",
                   threads_tried=pick_threads(g)[1])
        dt = float(np.abs(gpu_poses[:, :3] - gc.poses[:, :3]).max())
        rel = abs(sp.final_cost - r["final_cost"]) / r["final_cost"]
        parity = dict(solve="to termination, ftol 1e-6 / gtol 1e-10 / ptol 1e-8, 300 max, non-monotonic", n_gpus=world,
                      gpu_termination=sp.termination, cpu_termination=r["termination"],
                      gpu_lm_iterations=sp.num_lm_steps, cpu_lm_iterations=r["lm_steps"],
                      gpu_final_cost=sp.final_cost, cpu_final_cost=r["final_cost"], final_cost_rel_diff=rel,
                      max_pose_translation_diff_m=dt, bar="final cost 1e-5 relative, translations 1e-4 m (BASELINE.json north_star)",
                      ok=bool(sp.termination == r["termination"] and rel <= 1e-5 and dt <= 1e-4))
        nparam = (g.poses.size + g.points.size + g.objects.size) * 8
        line = dict(metric=METRIC, value=steps / dev_t, unit=UNIT, n_gpus=world, steps=steps, warmup=args.warmup,
                    ms_per_step=1e3 * dev_t / steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                    data="synthetic", config=dict(workload(args), parallelism=f"e-blocks sharded over {world} rank(s), reduced system all-reduced" if world > 1 else "single GPU",
                                                  counts=g.counts(), seed=args.seed),
                    clocks=clocks,
                    e2e=dict(value=steps / wall_t, unit=UNIT, h2d_bytes_per_step=nparam / steps, d2h_bytes_per_step=nparam / steps + 16 * 8,
                             note="obvi_solve through the C ABI with host parameter blocks: gather + H2D of every block, K iterations "
                                  "(each reads its 16-double result block -- cost, model change, norms, flags -- back through pinned "
                                  "memory and decides accept / reject on the host), D2H + scatter of every block; factor records are "
                                  "device-resident from the first solve, as in a persistent ceres::Problem",
                             first_call_preprocess_s=s_first.preprocessor_time_in_seconds),
                    e2e_cold=dict(value=steps / (s_first.preprocessor_time_in_seconds + wall_t), unit=UNIT,
                                  structure_build_and_upload_s=s_first.preprocessor_time_in_seconds,
                                  first_call_wall_s=t_first, first_call_lm_iterations=s_first.num_lm_steps,
                                  note="a one-shot user: the structure build + upload of the FIRST obvi_solve on a fresh problem added to the "
                                       "timed K-iteration call (the reference's analogue is its per-solve problem build, optimizer_build_pgo)"),
                    gpu_launches=int(s.kernel_launches), roofline=roof, cpu_baseline=cpu, parity=parity,
                    final_cost=gpu_cost, pcg_iterations=int(s.pcg_iterations_total),
                    phases_ms_per_step=dict(jacobian=1e3 * s.jacobian_evaluation_time_in_seconds / steps,
                                            linear=1e3 * s.linear_solver_time_in_seconds / steps,
                                            residual=1e3 * s.residual_evaluation_time_in_seconds / steps))
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    claim_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
