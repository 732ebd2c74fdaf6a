#!/usr/bin/env python
"""Headline benchmark: shooting-nodes/sec (forward + Jacobian/Hessian blocks) of the quadruped NMPC sweep.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...     the CPU implementation of the same path

One "step" = one pass of the hot path over one batch of synthetic trajectories: the KKT sweep
(every value / Jacobian block / Gauss-Newton Hessian block of every shooting node), the per-trajectory
summary kernel and, for N > 1 ranks, one all-gather of the summaries.  Workload at N = 1:
BASELINE.json configs[3] — quadruped SRBD, horizon 100, 1024 trajectories, fp64.  For N > 1 every rank keeps
1024 trajectories (weak scaling; 8 ranks = the 8192 trajectories of configs[4]).

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for how every field is measured.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from ungar_b200 import sharding  # noqa: E402
from ungar_b200 import workloads as W  # noqa: E402

METRIC = "shooting_nodes_per_sec_fwd_jac_quadruped_nmpc_N100"
UNIT = "nodes/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent
BARRIER = {W.QUADROTOR: (100.0, 2e-5), W.RC_CAR: (100.0, 1e-2), W.QUADRUPED: (1.0, 1.0)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--model", default="quadruped", choices=list(W.MODEL_IDS))
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--batch", type=int, default=1024, help="trajectories per GPU (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: this many trajectories in total, split evenly over the ranks (BASELINE.json configs[4]: 8192)")
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short legs on BASELINE configs 2 and 3 (quadrotor / RC car, fp32) that the default line carries as other_configs")
    ap.add_argument("--mode", default="kkt", choices=["kkt", "jacobian"],
                    help="kkt: full KKT block set (default, BASELINE configs[2..4]); jacobian: g + A only (configs[1], ungar_b200_jacobian_blocks)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# Workload description shared by both arms
# ------------------------------------------------------------------------------------------------------
BASELINE_CONFIG = {("quadrotor", "jacobian"): 1, ("quadrotor", "kkt"): 1, ("rc_car", "kkt"): 2, ("quadruped", "kkt"): 3}


def workload_config(args, world: int, record_bytes: int = 0, input_bytes: int = 0) -> dict:
    what = "KKT sweep" if args.mode == "kkt" else "Jacobian sweep (g + A only)"
    cfg = BASELINE_CONFIG.get((args.model, args.mode))
    return {
        "workload": f"{args.model} NMPC {what}, N={args.horizon}, {args.batch} trajectories/GPU, {args.dtype} "
                    f"({'BASELINE.json configs[%d]' % cfg if cfg is not None else 'not a BASELINE config'}; {world} GPU(s) -> "
                    f"{args.batch * world} trajectories)",
        "model_problem": args.model, "horizon": args.horizon, "batch_per_gpu": args.batch, "mode": args.mode,
        "global_batch": args.batch * world, "parallelism": f"independent trajectories sharded over {world} rank(s)",
        "scaling": "strong" if args.global_batch else "weak",
        "l2_policy": f"no flush: each step streams {record_bytes / 1e9:.2f} GB of records through the 126 MB L2 (evicting whatever the "
                     f"previous step left there) and the inputs rotate over 4 buffers ({4 * input_bytes / 1e6:.0f} MB)",
    }


def algorithmic_bytes_per_trajectory(layout: dict, elem: int, mode: str = "kkt") -> int:
    """every input scalar read once + every output scalar written once (padding excluded;
    quadruped contact rows in their compact 4x20 form, 4x10 for k = 0).  mode "jacobian": inputs + g + A (+ C) only (config 2)."""
    L = layout
    N = L["horizon"]
    contact = N * L["legs"] * 80 - L["legs"] * 40 if L["legs"] else 0
    if mode == "jacobian":
        return (L["n_dec"] + L["n_par"] + L["m_eq"] + N * L["nx"] * L["nz"] + contact) * elem
    scalars = (L["n_dec"] + L["n_par"] + L["m_eq"] + N * L["nx"] * L["nz"] + contact + L["m_ineq"] + 2 + L["n_dec"] +
               N * L["tri"] + L["tri_terminal"] + (N - 1) * L["hc_per_node"])
    return scalars * elem


def compact_algorithmic_bytes_per_trajectory(layout: dict) -> int:
    """Compact record (csrc/compact.cuh): inputs read once + the structurally non-zero outputs written once.  Per node: A 238 (the
    reference's CppAD pattern, function.hpp:98-105; the layout keeps 12 more slots that are structural zeros), H 61, C 224
    (k = 0: 128), g 29, h 12, q 37; per trajectory: x_0 - x_m 13, q_N 13, H_N diagonal 13, cost 2.  Pads excluded."""
    N = layout["horizon"]
    per_node = 238 + 61 + 224 + 29 + 12 + 37
    return (layout["n_dec"] + layout["n_par"] + N * per_node - 96 + 13 + 13 + 13 + 2) * 8


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return FALLBACK_HBM_GBS, "B200_PROFILING.md fallback (of fallback)"


def recorded_traffic(key: str):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------
# Clock sampling during the timed region (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.windows = []
        self._proc = None
        self._thread = None
        self.index = index

    def start(self):
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._proc = None
            return
        self._thread = threading.Thread(target=self._pump, daemon=True)
        self._thread.start()

    def _pump(self):
        for line in self._proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                try:
                    self.samples.append((time.time(), float(parts[0]), float(parts[1]), parts[3:7]))
                except ValueError:
                    pass

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self) -> dict:
        if self._proc is not None:
            self._proc.terminate()  # the exact PID we started
            try:
                self._proc.wait(timeout=5)
            except Exception:
                self._proc.kill()
        inside = [s for s in self.samples if any(a - 0.05 <= s[0] <= b + 0.05 for a, b in self.windows)]
        use = inside or self.samples
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in use for n, v in zip(names, s[3]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(s[1] for s in use), "sm_max_mhz": max(s[2] for s in use),
                "reasons": reasons, "samples": len(use), "samples_in_timed_region": len(inside)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's stage-wise port (kind "port"; the reference itself cannot be built: DESIGN.md §4)
# ------------------------------------------------------------------------------------------------------
def cpu_sweep_runner(args):
    """Returns (run(sample_xp) -> (seconds, records), threads).  ORACLE USE: timed CPU baseline only."""
    import oracle

    oracle.build()
    orc = oracle.Oracle(fast=True)
    threads = os.cpu_count() or 1
    mid = W.MODEL_IDS[args.model]
    k, eps = BARRIER[mid]
    size = orc.record_layout(mid, args.horizon)["size"]

    def run(xp):
        out = np.empty((xp.shape[0], size))
        t0 = time.perf_counter()
        orc.stage_sweep(mid, args.horizon, xp, k, eps, threads=threads, out=out)
        return time.perf_counter() - t0, out

    return run, threads


def cpu_sample_size(run, threads, xp_pool, seconds):
    probe = xp_pool[:min(len(xp_pool), 2 * threads)]
    run(probe)  # warm caches / page in
    t, _ = run(probe)
    per_traj = t / len(probe)
    n = int(max(threads, min(len(xp_pool), seconds / max(per_traj, 1e-9))))
    return max(threads, (n // threads) * threads)


def reference_arm(args, rank: int, world: int):
    """The reference's CPU implementation of the path on all host threads.  Two stand-ins exist for the absent CppADCodeGen binary: the
    generated straight-line code (oracle/codegen_baseline.py, kind "codegen": what the reference's MakeFunction would produce) and the
    hand-written dense stage-wise port (oracle/stage_port.cpp, kind "port").  The faster of the two is the arm's value."""
    if rank != 0:
        return  # the CPU arm has no multi-process path: rank 0 alone runs and prints it
    mid = W.MODEL_IDS[args.model]
    run, threads = cpu_sweep_runner(args)
    import oracle as oracle_pkg  # the CPU arm may execute oracle/ (it IS the thing timed here)

    ref_record_size = oracle_pkg.Oracle().record_layout(mid, args.horizon)["size"]  # same layout as ungar_b200_kkt_layout
    pool = W.synthetic_batch(mid, args.horizon, min(args.batch, 1024), seed=20240807)
    total_steps = args.steps + args.warmup
    per_step = min(60.0 / max(total_steps, 1), 10.0)  # whole run within ~1-2 minutes
    # pick the faster stand-in on a short probe
    kind, cg = "port", None
    try:
        from oracle import codegen_baseline as CG

        cg = CG.Baseline(args.model, args.horizon)
        probe = pool[:2 * threads]
        cg.run(probe, threads)
        t_cg, _ = cg.run(probe, threads)
        run(probe)
        t_port, _ = run(probe)
        if t_cg < t_port:
            kind = "codegen"
    except Exception:
        cg = None
    if kind == "codegen":
        buf = {}

        def run_ref(xp):
            if buf.get("n") != xp.shape[0]:
                buf["out"], buf["n"] = np.empty((xp.shape[0], cg.total)), xp.shape[0]
            return cg.run(xp, threads, buf["out"])
    else:
        run_ref = run
    n = cpu_sample_size(run_ref, threads, pool, per_step)
    sample = pool[:n]
    for _ in range(args.warmup):
        run_ref(sample)
    t_total = 0.0
    for _ in range(args.steps):
        t, _ = run_ref(sample)
        t_total += t
    value = n * args.horizon * args.steps / t_total
    what = (f"oracle/codegen_baseline.py: straight-line C from the reference's tapes, gcc -O3 -ffast-math -march={'native' if cg and cg.variant == 'native' else 'x86-64-v3'}"
            if kind == "codegen" else "oracle/stage_port.cpp built -O3 -ffast-math -march=x86-64-v3")
    desc = f"{n} of the workload's {args.batch} trajectories per step, {args.steps} steps, fp64, {what}"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, max(args.gpus, 1), args.batch * ref_record_size * 8, args.batch * W.sizes(mid, args.horizon)["n_xp"] * 8),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU arm: the reference's CppAD/CppADCodeGen path cannot be built offline; this is the faster of the two stand-ins "
                "(generated straight-line code from the reference's tapes / the oracle's stage-wise port) on all host threads (rank 0 only).",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# Our arm
# ------------------------------------------------------------------------------------------------------
def ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    import ungar_b200
    from ungar_b200 import parity as P

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mid = W.MODEL_IDS[args.model]
    if args.global_batch:  # strong scaling: the total is fixed, every rank takes an equal contiguous shard
        lo, hi = sharding.shard_range(args.global_batch, world, rank)
        args.batch = hi - lo
    B, N = args.batch, args.horizon
    model = ungar_b200.Model(args.model, N, dtype=args.dtype, device=local_rank, barrier=BARRIER[mid])
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    elem = 8 if args.dtype == "f64" else 4
    L = model.layout
    jac = args.mode == "jacobian"
    has_compact = args.model == "quadruped" and args.dtype == "f64" and not jac
    cmodel = ungar_b200.Model(args.model, N, dtype=args.dtype, device=local_rank, barrier=BARRIER[mid], record_format="compact") if has_compact else None

    # synthetic inputs: host copy in pinned memory (e2e leg), 4 rotating device copies (device-resident leg)
    xp_np = W.synthetic_batch(mid, N, B, seed=20240807 + rank).astype(model.np_dtype)
    xp_host = torch.empty((B, model.n_xp), dtype=tdt, pin_memory=True)
    xp_host.copy_(torch.from_numpy(xp_np))
    d_xps = [xp_host.to(dev, non_blocking=True).clone() for _ in range(4)]
    d_rec = torch.empty((B, L["size"]), dtype=tdt, device=dev)
    d_sum = torch.empty((B, 32), dtype=tdt, device=dev)
    gathered = torch.empty((world * B, 32), dtype=tdt, device=dev) if world > 1 else None
    sum_host = torch.empty((B, 32), dtype=tdt, pin_memory=True)

    exchange = sharding.SummaryExchange(B, tdt, dev) if world > 1 and not jac else None

    def step(i, mdl=model, rec=d_rec):
        # the all-gather of step i is posted asynchronously and overlaps the sweep of step i + 1 (double-buffered summaries)
        if jac:
            mdl.jacobian_blocks(d_xps[i % 4], rec)
            return
        if exchange is None:
            mdl.step(d_xps[i % 4], records=rec, summaries=d_sum)
            return
        k = exchange.slot()
        mdl.step(d_xps[i % 4], records=rec, summaries=exchange.local[k])
        exchange.post(k)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_steps(fn, steps, warmup):
        """W warm-up + K timed calls of fn(i), bracketed by barrier + synchronize; returns (max-over-ranks ms, sweep ms list, launches)."""
        for i in range(warmup):
            fn(i)
        fence()
        model.set_profiling(True)
        launches0 = model.launch_count()
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        if exchange is not None:
            exchange.drain()  # the last gathers are inside the timed region
        e1.record()
        fence()
        t1 = time.time()
        launches = model.launch_count() - launches0
        sweep_ms = model.sweep_times_ms()
        model.set_profiling(False)
        if sampler:
            sampler.window(t0, t1)
        return max_over_ranks(e0.elapsed_time(e1)), sweep_ms, launches

    # ---- device-resident leg (the handle's DENSE records: the byte count SURVEY.md 8d defines) ---------------------------------
    elapsed_ms, sweep_ms, launches = timed_steps(step, args.steps, max(args.warmup, 3))
    value = world * B * N * args.steps / (elapsed_ms * 1e-3)

    # ---- the same sweep writing COMPACT records (only the structurally non-zero slots) -----------------------
    compact = None
    if has_compact:
        c_rec = torch.empty((B, cmodel.layout["size"]), dtype=tdt, device=dev)
        c_ms, c_sweep_ms, _ = timed_steps(lambda i: step(i, cmodel, c_rec), args.steps, max(args.warmup, 3))
        c_bytes = compact_algorithmic_bytes_per_trajectory(L) * B
        c_kernel_ms = statistics.fmean(c_sweep_ms) if c_sweep_ms else c_ms / args.steps
        peak_c, _ = measured_peak_gbs()
        compact = {"value": world * B * N * args.steps / (c_ms * 1e-3), "unit": UNIT, "ms_per_step": c_ms / args.steps,
                   "record_bytes_per_trajectory": cmodel.layout["size"] * elem, "dense_record_bytes_per_trajectory": L["size"] * elem,
                   "roofline": {"bound": "hbm", "achieved": c_bytes / (c_kernel_ms * 1e-3) / 1e9, "peak": peak_c, "unit": "GB/s",
                                "frac": c_bytes / (c_kernel_ms * 1e-3) / 1e9 / peak_c, "kernel": "quadruped_compact_kernel",
                                "kernel_ms": c_kernel_ms, "algorithmic_bytes_per_launch": c_bytes,
                                "traffic": recorded_traffic(f"{args.model}_{args.dtype}_N{N}_B{B}_compact")},
                   "path": "ungar_b200_kkt_step on a RECORD_COMPACT handle: same arithmetic, one 5008-byte bulk store per node"}

    # ---- end-to-end leg: host buffers through the C ABI.  The parameter block (references, constants, measured state) is cached on the
    # device once per control cycle (ungar_b200_set_parameters, outside the timed region); every step copies the DECISION variables
    # host -> device (pinned memory), sweeps, and reads the summaries back.  `e2e_full_xp` is round 1's variant (whole xp every step).
    e2e_steps = max(10, min(args.steps, 100))
    xp_host_np, sum_host_np = xp_host.numpy(), sum_host.numpy()
    rec_host_np = torch.empty((B, L["size"]), dtype=tdt, pin_memory=True).numpy() if jac else None
    x_host = torch.empty((B, L["n_dec"]), dtype=tdt, pin_memory=True)
    x_host.copy_(xp_host[:, :L["n_dec"]])
    x_host_np = x_host.numpy()
    if not jac:
        model.set_parameters(np.ascontiguousarray(xp_np[:, L["n_dec"]:]))

    def e2e_gather():
        if world > 1:  # the step's own summaries (they landed on the host) are what the ranks exchange
            d_sum.copy_(sum_host, non_blocking=True)
            dist.all_gather_into_tensor(gathered, d_sum)

    def e2e_step(_i):
        if jac:  # the ABI returns the record to the host (g and A valid): H2D of xp, sweep, D2H of the record
            model.jacobian_blocks(xp_host_np, rec_host_np)
            return
        model.step_x(x_host_np, records=d_rec, summaries=sum_host_np)
        e2e_gather()

    def e2e_step_full_xp(_i):
        model.step(xp_host_np, records=d_rec, summaries=sum_host_np)
        e2e_gather()

    def timed_host(fn, steps):
        for i in range(3):
            fn(i)
        fence()
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        fence()
        if sampler:
            sampler.window(t0, time.time())
        return max_over_ranks(e0.elapsed_time(e1))

    if jac:
        e2e_steps = min(e2e_steps, 20)
    e2e_ms = timed_host(e2e_step, e2e_steps)
    e2e_value = world * B * N * e2e_steps / (e2e_ms * 1e-3)
    e2e_full_xp = None
    if not jac:
        fx_ms = timed_host(e2e_step_full_xp, e2e_steps)
        e2e_full_xp = {"value": world * B * N * e2e_steps / (fx_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * model.n_xp * elem,
                       "path": "ungar_b200_kkt_step(MEM_HOST): the whole flat vector [X | U | parameters] crosses PCIe every step (round 1's e2e)"}

    # the collective alone (report it separately): one all-gather of the [B, 32] summaries per outer iteration
    collective = None
    if world > 1:
        fence()
        e0.record()
        for _ in range(args.steps):
            dist.all_gather_into_tensor(gathered, d_sum)
        e1.record()
        fence()
        coll_ms = max_over_ranks(e0.elapsed_time(e1))
        collective = {"op": "all_gather_into_tensor (NCCL)", "bytes_per_rank": B * 32 * elem, "ms_per_call": coll_ms / args.steps,
                      "share_of_step": coll_ms / elapsed_ms}

    # full drop-in variant: the whole record goes back to the host every step (what a host-side QP solver needs); compact records
    full_value = None
    if world == 1 and not jac:
        try:
            fm = cmodel if has_compact else model
            rec_host = torch.empty((B, fm.layout["size"]), dtype=tdt, pin_memory=True).numpy()
            fm.kkt_blocks(xp_host_np, rec_host)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                fm.kkt_blocks(xp_host_np, rec_host)
            full_value = {"value": B * N * reps / (time.perf_counter() - t0), "unit": UNIT, "d2h_bytes_per_step": B * fm.layout["size"] * elem,
                          "record_format": "compact" if has_compact else "dense",
                          "path": "ungar_b200_kkt_blocks(MEM_HOST): whole record back to the host every step"}
        except Exception:
            full_value = None

    # ---- the loop that consumes the records: sweep + QP solve + line search on the device.  Every rank solves its
    # shard; per outer iteration the ranks all-gather the POST-SOLVE summaries (u_0*, f*, |g|, max h at the solution): north_star's
    # "one NCCL all-gather of the optimal-control summaries per outer iteration".
    sqp = None
    if args.dtype == "f64" and not jac:
        try:
            iters = 4 if args.model == "quadruped" else 10  # quadruped.example.cpp:444 (4); the other two examples allow 40 and stop early
            mult = 1.0 if args.model == "quadrotor" else 1.0 / N
            opts = model.sqp_options(max_iterations=iters, constraint_violation_multiplier=mult)
            work = d_xps[0].clone()
            sol_sum = torch.empty((B, 32), dtype=tdt, device=dev)

            def outer_iteration():
                work.copy_(d_xps[0])
                status, _ = model.sqp_solve(work, opts, want_info=False)
                model.step(work, records=d_rec, summaries=sol_sum)  # summaries AT the solution: u_0*, f*, |g|_inf, max h
                if world > 1:
                    dist.all_gather_into_tensor(gathered, sol_sum)
                return status

            outer_iteration()
            fence()
            reps = 3
            e0.record()
            for _ in range(reps):
                status = outer_iteration()
            e1.record()
            fence()
            ms = max_over_ranks(e0.elapsed_time(e1)) / reps
            # the solve alone on this rank (no copy-in, no summaries, no collective)
            work.copy_(d_xps[0])
            torch.cuda.synchronize()
            e0.record()
            model.sqp_solve(work, opts, want_info=False)
            e1.record()
            torch.cuda.synchronize()
            solve_ms = e0.elapsed_time(e1)
            counts = torch.bincount(status[:, 0], minlength=3).tolist()
            sqp = {"ms_per_solve": solve_ms, "ms_per_outer_iteration": ms, "iterations": iters, "trajectories": world * B,
                   "trajectory_iterations_per_sec": world * B * iters / (ms * 1e-3),
                   "status_counts_rank0": {"max_iterations": counts[0], "converged": counts[1], "line_search_failed": counts[2]},
                   "path": "per rank: ungar_b200_sqp_solve(MEM_DEVICE) = iterations x {KKT sweep (compact records), QP solve, backtracking line "
                           "search}, then one sweep at the solution for the summaries" + (", then one NCCL all-gather of the [B, 32] post-solve summaries" if world > 1 else "")}
            # end to end through host buffers: upload xp once, solve, download the solution and the statuses
            # (pinned buffer, updated in place by the solve: restoring the initial guess between repetitions is not part of a solve)
            h_work = torch.empty((B, model.n_xp), dtype=tdt, pin_memory=True).numpy()
            h_work[:] = xp_host_np
            model.sqp_solve(h_work, opts, want_info=False)
            fence()
            total = 0.0
            for _ in range(reps):
                h_work[:] = xp_host_np
                t0 = time.perf_counter()
                model.sqp_solve(h_work, opts, want_info=False)  # returns after the stream has drained (cudaStreamSynchronize)
                total += time.perf_counter() - t0
            host_ms = total / reps * 1e3
            host_ms = max_over_ranks(host_ms)
            sqp["e2e"] = {"ms_per_solve": host_ms, "trajectory_iterations_per_sec": world * B * iters / (host_ms * 1e-3),
                          "h2d_bytes_per_solve": B * model.n_xp * elem, "d2h_bytes_per_solve": B * (L["n_dec"] * elem + 8),
                          "path": "ungar_b200_sqp_solve(MEM_HOST): xp host -> device once, the whole loop on the device, solution + statuses back"}
        except Exception as exc:  # auxiliary figure: never fails the bench line
            sqp = {"error": str(exc)}
        if world == 1 and sqp and "error" not in sqp and not args.no_cpu_baseline:
            sqp["cpu_baseline"] = sqp_cpu_baseline(xp_np, N, BARRIER[mid], sqp["iterations"], mult, model=mid)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the sweep), from device events around each launch ---------------
    bytes_per_launch = algorithmic_bytes_per_trajectory(L, elem, args.mode) * B
    peak, peak_src = measured_peak_gbs()
    mean_sweep_ms = statistics.fmean(sweep_ms) if sweep_ms else elapsed_ms / args.steps
    achieved = bytes_per_launch / (mean_sweep_ms * 1e-3) / 1e9
    key = f"{args.model}_{args.dtype}_N{N}_B{B}" + ("_jacobian" if jac else "")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(key), "kernel": "kkt_sweep", "kernel_ms": mean_sweep_ms,
                "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
                "kernel_share_of_step": mean_sweep_ms / (elapsed_ms / args.steps)}

    # ---- parity gate on rank 0 at every N (strict metric of SURVEY.md 8d, ungar_b200/parity.py) + CPU baseline at N = 1 ---------
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        run, threads = cpu_sweep_runner(args)
        pool = xp_np.astype(np.float64)
        if world == 1:
            n = cpu_sample_size(run, threads, pool, args.cpu_seconds)
            t, ref = run(pool[:n])
            passes, t_total = 1, t
            while t_total < args.cpu_seconds and passes < 10000:  # ~10-30 s of CPU work on the bounded sample
                t, ref = run(pool[:n])
                t_total += t
                passes += 1
            cpu = {"value": n * N * passes / t_total, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"first {n} of the {B} trajectories x {passes} passes, fp64, oracle/stage_port.cpp "
                             f"(-O3 -ffast-math) on {threads} threads ({t_total:.1f} s of wall time)"}
            codegen = cpu_codegen_baseline(args, pool)
            if codegen:
                cpu["codegen"] = codegen
        else:  # N > 1: the oracle only as the checker, on this rank's first trajectories
            n = min(B, 256)
            _, ref = run(pool[:n])
        got = d_rec_sample(model, d_xps[0], n, tdt, dev, jac)
        keys = ("g", "A") + (("C",) if L["legs"] else ()) if jac else None  # only g and A are specified after the Jacobian sweep
        rep = P.compare_records(model.split_record, got, ref, xp_np[:n], L["nx"] * (N + 1), args.dtype, keys=keys)
        parity = {"checked_trajectories": n, "tolerance": P.RTOL[args.dtype], **rep}
        if has_compact:  # dense-from-compact against the same oracle records
            gotc = d_rec_sample(cmodel, d_xps[0], n, tdt, dev, False)
            repc = P.compare_records(model.split_record, cmodel.to_dense(gotc), ref, xp_np[:n], L["nx"] * (N + 1), args.dtype)
            parity["compact"] = {k: repc[k] for k in ("ok", "max_rel_err", "strict_failures", "cancellation_entries", "legacy_max_rel_err")}
            parity["ok"] = bool(parity["ok"] and repc["ok"])
        if not parity["ok"]:
            print(json.dumps({"error": "parity gate failed", "parity": parity}), flush=True)
            raise SystemExit(2)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, world, B * L["size"] * elem, B * model.n_xp * elem), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (model.n_xp if jac else L["n_dec"]) * elem,
                "d2h_bytes_per_step": B * (L["size"] if jac else 32) * elem, "steps": e2e_steps,
                "path": "ungar_b200_jacobian_blocks(MEM_HOST): pinned host xp -> H2D -> Jacobian sweep -> record -> D2H" if jac else
                        "ungar_b200_kkt_step_x(MEM_HOST): pinned host decision variables [X | U] -> H2D in 8 linear chunks (landing zone, scattered on the device) overlapped with the sweep of the previous "
                        "chunk (parameter block cached on the device by ungar_b200_set_parameters once per control cycle; records stay in HBM) -> summaries -> D2H"},
        "e2e_full_xp": e2e_full_xp, "e2e_full_record_d2h": full_value, "compact": compact,
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "sqp_loop": sqp,
        "collective": collective,
    }
    if default_headline(args) and world == 1:
        line["other_configs"] = other_configs(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def default_headline(args) -> bool:
    return (args.model == "quadruped" and args.horizon == 100 and args.dtype == "f64" and args.mode == "kkt" and not args.global_batch
            and not args.no_other_configs and not args.no_cpu_baseline)


def other_configs(args) -> dict:
    """BASELINE.json configs 2 and 3 (the fp32 small-model sweeps north_star's roofline clause is about) measured by this same script
    in short sub-runs, so that their numbers are in the driver's bench line and not only in builder-run records: device-resident
    value, kernel time, roofline fraction on the config's own byte count, strict parity against the oracle, end-to-end value.
    Each sub-run is `python bench.py --model ... --no-other-configs` (its own process: a model handle per process keeps the timing clean)."""
    out = {}
    runs = {
        "configs[1] quadrotor N=30 B=4096 f32, Jacobian sweep (g + A)": ["--model", "quadrotor", "--horizon", "30", "--batch", "4096", "--dtype", "f32", "--mode", "jacobian"],
        "configs[1] quadrotor N=30 B=4096 f32, full KKT block set": ["--model", "quadrotor", "--horizon", "30", "--batch", "4096", "--dtype", "f32"],
        "configs[2] rc_car N=60 B=8192 f32, full KKT block set": ["--model", "rc_car", "--horizon", "60", "--batch", "8192", "--dtype", "f32"],
    }
    for name, extra in runs.items():
        cmd = [sys.executable, os.path.abspath(__file__), "--steps", "200", "--warmup", "10", "--cpu-seconds", "2", "--no-other-configs"] + extra
        try:
            proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            d = json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1])
            if "error" in d:
                out[name] = d
                continue
            par = d.get("parity") or {}
            out[name] = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "dtype": d["dtype"],
                         "roofline": {k: d["roofline"].get(k) for k in ("achieved", "peak", "unit", "frac", "kernel_ms", "algorithmic_bytes_per_launch")},
                         "e2e": d["e2e"]["value"], "cpu_baseline": (d.get("cpu_baseline") or {}).get("value"),
                         "parity": {k: par.get(k) for k in ("ok", "tolerance", "max_rel_err", "strict_failures", "cancellation_entries",
                                                            "cancellation_worst_ulps", "checked_trajectories")},
                         "command": "python bench.py " + " ".join(extra)}
        except Exception as exc:  # an auxiliary figure must not fail the headline line
            out[name] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def cpu_codegen_baseline(args, pool):
    """The CPU baseline SURVEY.md 8d specifies: straight-line C generated from the reference's own tapes, gcc -O3 -march=native
    -ffast-math (function.hpp:610-611), timed at 1 thread and on all cores (oracle/codegen_baseline.py).  None when unavailable."""
    try:
        from oracle import codegen_baseline as CG

        return CG.timed(args.model, args.horizon, pool, seconds=min(args.cpu_seconds, 8.0))
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def sqp_cpu_baseline(xp_np, N, barrier, iterations, multiplier, seconds=8.0, model=W.QUADRUPED):
    """The same soft-SQP solve on the host cores: oracle/sqp_port.cpp (stage sweep -> exact stage-wise QP -> line search), all
    threads, on a bounded sample of the bench's trajectories.  Never raises: an auxiliary figure must not fail the bench line."""
    try:
        import oracle as oracle_pkg

        orc = oracle_pkg.Oracle(fast=True)
        threads = os.cpu_count() or 1
        pool = np.ascontiguousarray(xp_np, dtype=np.float64)
        n = min(len(pool), threads)
        t0 = time.perf_counter()
        orc.sqp_solve_port(N, pool[:n], barrier[0], barrier[1], multiplier, iterations, threads=threads, model=model)
        per_round = max(time.perf_counter() - t0, 1e-6)  # one trajectory per thread
        n = int(min(len(pool), max(threads, threads * int(seconds / per_round))))
        t0 = time.perf_counter()
        _, status = orc.sqp_solve_port(N, pool[:n], barrier[0], barrier[1], multiplier, iterations, threads=threads, model=model)
        dt = time.perf_counter() - t0
        done = int(status[:, 1].sum())
        return {"value": done / dt, "unit": "trajectory-iterations/s", "cores": threads, "kind": "port",
                "sample": f"first {n} trajectories x {iterations} iterations, fp64, oracle/sqp_port.cpp (-O3 -ffast-math: stage sweep, dense "
                          f"stage-wise QP (Schur complement / Riccati), backtracking line search) on {threads} threads ({dt:.1f} s); the reference itself hands the "
                          f"QP to OSQP (ADMM), absent here"}
    except Exception as exc:
        return {"error": str(exc)}


def d_rec_sample(model, d_xp, n, tdt, dev, jac=False):
    import torch

    out = torch.zeros((n, model.layout["size"]), dtype=tdt, device=dev)
    (model.jacobian_blocks if jac else model.kkt_blocks)(d_xp[:n], out)
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64)


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.gpus > 1 and world == 1:  # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
