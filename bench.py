#!/usr/bin/env python
"""bench.py — psi evaluations/sec (subject x support-point pairs) of the B200 psi-matrix backend.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...    # the path's CPU implementation on host cores

A "step" is one full psi matrix: every (subject, support point) pair of the workload simulated and
its log-likelihood stored (column-major), with the population, the support points and the output
resident in HBM.  Headline workload = BASELINE.json configs[1]: two-compartment oral absorption
`ode!` (Dopri5), 500 subjects x 20,000 support points, 10 doses + 12 observations per subject.
For N > 1 (torchrun, one rank per GPU) every rank owns 20,000 support-point columns (weak scaling)
and the step ends with the gather of the psi column slabs onto every rank.

Besides the headline the default run appends
  * `configs`  (N = 1): the other BASELINE configs — C1, C3 (one GPU's shard of the 8-GPU config), C4 (with the
    model forced through NVRTC, compile time reported) and C5 (one step) — each with value / e2e / roofline /
    cpu_baseline / psi_nan, so every BASELINE config has a driver-run record;
  * `strong`   (every N): STRONG scaling — C2 at 20,000 columns in total and C3 at its full BASELINE size
    10,000 x 50,000 in total, split over the N ranks, with the kernel and gather shares per step;
  * `gather_matches_single_gpu` (N > 1): every rank recomputes a block of ANOTHER rank's columns on its own GPU and
    compares it bit for bit with what arrived in its copy of psi; the run fails when that is false;
  * `library_multi_device` (N > 1, rank 0): the same matrix through ONE process and the C ABI's multi-device
    context (pharmsol_cuda_ctx_create_multi), host buffers in and out.

Timing: W >= 3 untimed warm-up steps; K timed steps, each bracketed by CUDA events on the launch
stream; an L2 flush (a 512 MiB device memset, outside the event bracket) separates iterations;
barrier + synchronize on both sides of the timed region; max over ranks.  `e2e` times the same
step through the public host-buffer API (H2D of the support points + D2H of psi inside the region)
with pinned caller buffers; `e2e_pageable` with plain malloc'ed numpy arrays (what a Rust `Array2::zeros` is).
`roofline` is FP64 (the path is compute-bound FP64 scalar work): algorithmic flops are
counted from the device's own step / RHS counters times the per-step figures of DESIGN.md, the peak
is the DFMA-chain throughput measured in this run.  `cpu_baseline` times the oracle restatement
(OpenMP over subjects == the reference's rayon decomposition) on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "psi evaluations/sec (subject x support-point pairs)"
UNIT = "pairs/s"

# workload -> (solver name, algorithmic flop model)  — see DESIGN.md "Algorithmic work"
WORKLOADS = {
    "c1": dict(nsub=1000, nspp=1000, solver=None),
    "c2": dict(nsub=500, nspp=20000, solver="Dopri5"),
    "c3": dict(nsub=10000, nspp=6250, solver=None),        # 50k columns / 8 GPUs = 6,250 per GPU
    "c4": dict(nsub=2000, nspp=10000, solver="Rodas4"),
    "c5": dict(nsub=200, nspp=5000, solver=None),
}
# SURVEY §8d op weights: add/sub/mul = 1, fma = 2, div = sqrt = 10, exp = log = 24, sincos = 40 each, atan2 = 50, pow = 60
F_RHS = {"c2": 10.0, "c4": 2 * 10 + 8.0}
# Dopri5 / Tsit5 step.  SURVEY's convention counts 2n(21 + 7 + 7) + 60: 21 a-coefficients, the b row and the error row.
# The b row IS row 7 of A (FSAL) and one a-coefficient is zero, so the instruction-true figure is 2n(20 + 6 + ...) — both
# are reported: `frac` follows the survey's convention, `frac_instruction_true` the flop-true count.
C5_FLOP_PER_EVAL = 14.0
# warp instructions issued per particle attempt of the SDE kernel (ncu, profiles/r02_final_c5_ncu.txt: smsp__inst_executed
# 4.84e11 over 8.56e10 thread-attempts, 26.4 of 32 lanes active) — the kernel's real bound is instruction issue
C5_WARP_INST_PER_ATTEMPT = 5.66
ERK_STEP = lambda n: 2.0 * n * (21 + 7 + 7) + 60.0
ERK_STEP_TRUE = lambda n: 2.0 * n * (20 + 6 + 1) + 2.0 * n + 45.0       # stage sums + y + h*acc + error row + weights/controller


def algorithmic_flops(name, npairs, counters, nobs_per_subject, true_count=False):
    """Algorithmic FP64 work of one psi matrix (flop-equivalents, SURVEY §8d convention)."""
    if name == "c1":      # counters["evals"] = closed-form propagation steps executed on the device
        return counters["evals"] * (24 + 5) + npairs * (nobs_per_subject * 5 + 20)
    if name == "c3":
        return counters["evals"] * (850 + 75 + 4) + npairs * nobs_per_subject * 5
    if name == "c2":
        attempts = counters["steps"] + counters["rejected"]
        return counters["evals"] * F_RHS["c2"] + attempts * (ERK_STEP_TRUE(3) if true_count else ERK_STEP(3)) + npairs * nobs_per_subject * 15
    if name == "c4":
        # RODAS4 / SDIRK: per RHS F_rhs; per Newton iteration one 2x2 solve; per step one Jacobian + LU + stage combinations
        return counters["evals"] * F_RHS["c4"] + counters["newton"] * 30 + (counters["steps"] + counters["rejected"]) * 150
    if name == "c5":
        # ncu op counts of the reference stepper (profiles/r02_c5_source_attribution.txt): 14.2 DFMA + 11.0 DMUL + 3.0 DADD
        # per attempt = 42 flop; the device counts 3 evaluations per attempt
        return float(counters["evals"]) * C5_FLOP_PER_EVAL
    raise KeyError(name)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--nsub", type=int, default=0)
    ap.add_argument("--nspp", type=int, default=0, help="support points PER GPU (weak scaling) / in total with --scaling strong")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--tol", type=float, default=1e-6, help="ODE rtol = atol")
    ap.add_argument("--particles", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extras", default="auto", choices=["auto", "none"],
                    help="auto: the default run (headline C2) also records the other BASELINE configs (N = 1) and the strong-scaling points")
    ap.add_argument("--force-nvrtc", action="store_true", help="compile the model through NVRTC instead of the ahead-of-time twin")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "push", "nccl"],
                    help="N > 1: fused all-gather by peer stores from the psi kernel, copy-engine pushes of finished column chunks, "
                         "a separate NCCL all-gather, or auto (peer stores for ODE / SDE models, copy-engine pushes for closed-form "
                         "models whose psi is produced at GB/s rates)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_workload(name, nsub, nspp_total, particles):
    from benches import workloads as W
    kw = dict(nsub=nsub, nspp=nspp_total)
    if name == "c5":
        kw["particles"] = particles
    return W.make(name, **kw)


def config_dict(name, w, nsub, nspp_per_gpu, nspp_total, world, tol, particles, gather):
    cfg = {"workload": f"{name}: {w['desc']}", "nsub": nsub, "nspp_per_gpu": nspp_per_gpu, "nspp_total": nspp_total,
           "sharding": f"support-point columns x{world}",
           "l2": "512 MiB device memset between timed iterations (outside the event bracket)",
           "gather": gather}
    if w["kind"] == "ode":
        cfg.update(solver=WORKLOADS[name]["solver"], rtol=tol, atol=tol)
    if w["kind"] == "sde":
        cfg.update(particles=particles, sde_mode="particle filter (SDE::estimate_log_likelihood, sde/mod.rs:526-577)",
                   stepper="reference adaptive Euler-Maruyama (sde/em.rs:134-167)",
                   noise="standard normals by FP32 Box-Muller from Philox4x32-10 words (library default, DESIGN.md §4); state, drift, "
                         "diffusion, weights and likelihood arithmetic in FP64")
    return cfg


GATHER_TEXT = {
    None: "none (1 GPU)",
    "peer": "fused: psi kernel stores to every rank over NVLink + device barrier",
    "push": "copy engines: finished column chunks pushed to every rank over NVLink while the next chunk computes + device barrier",
    "nccl": "NCCL all_gather_into_tensor (in place)",
}


def gather_text(world, mode):
    return GATHER_TEXT[None] if world == 1 else GATHER_TEXT[mode]


def default_gather(kind, requested):
    """What `--gather auto` resolves to for a model kind (mirrors ResidentPsi)."""
    if requested != "auto":
        return requested
    return "push" if kind == "analytical" else "peer"


def observations_per_subject(w):
    ops = w["subjects"][0][1]
    return sum(1 for o in ops if o[0] == "observation")


# ---------------------------------------------------------------------------------------------------
def host_cores():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1: ask the OS, not OpenMP)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(name, w, nsub, tol, particles, budget_s, threads=0):
    """Time the oracle (restated CPU path, OpenMP over subjects) on a bounded sample: all subjects x
    the first S support points, S calibrated so the sample takes ~budget_s.  Returns (pairs/s, info)."""
    from benches import harness as H
    import oracle as O  # noqa: F401
    threads = threads or host_cores()
    kw = {}
    if w["kind"] == "ode":
        kw = dict(solver="dopri5", rtol=tol, atol=tol)
    if w["kind"] == "sde":
        kw = dict(particles=particles)
    om, od, oe = H.oracle_objects(w, **kw)
    spp = w["support_points"]
    s = min(len(spp), 64 if w["kind"] != "sde" else 1)
    t0 = time.perf_counter()
    mkw = dict(sde_mode=1) if w["kind"] == "sde" else {}
    om.log_likelihood_matrix(od, spp[:s], oe, nthreads=threads, **mkw)
    probe = time.perf_counter() - t0
    rate = nsub * s / max(probe, 1e-9)
    s2 = int(min(len(spp), max(s, rate * budget_s / nsub)))
    if w["kind"] == "sde":
        s2 = min(s2, 2)      # attempt counts vary strongly with (ke, sigma): keep the sample bounded
    t0 = time.perf_counter()
    _, info = om.log_likelihood_matrix(od, spp[:s2], oe, nthreads=threads, return_info=True, **mkw)
    dt = time.perf_counter() - t0
    return nsub * s2 / dt, {"cores": int(info["threads"]), "sample": f"{nsub} subjects x first {s2} support points of the workload, {dt:.1f} s",
                            "seconds": dt, "nspp_sample": s2}


CPU_NOTE = "restated CPU oracle (C++/OpenMP over subjects, the reference's rayon decomposition); the Rust reference cannot be built here (no cargo)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = WORKLOADS[args.workload]
    nsub = args.nsub or cfg["nsub"]
    nspp = args.nspp or cfg["nspp"]
    w = make_workload(args.workload, nsub, nspp, args.particles)
    total = args.steps + args.warmup
    budget = min(15.0, max(1.0, 150.0 / max(total, 1)))
    for _ in range(args.warmup):
        cpu_reference_rate(args.workload, w, nsub, args.tol, args.particles, budget)
    info = None
    t_all = 0.0
    pairs = 0
    for _ in range(args.steps):
        _, info = cpu_reference_rate(args.workload, w, nsub, args.tol, args.particles, budget)
        t_all += info["seconds"]
        pairs += nsub * info["nspp_sample"]
    value = pairs / t_all
    # the same `config` keys and values as the product arm's N = 1 line (the driver compares them)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, w, nsub, nspp, nspp, 1, args.tol, args.particles, gather_text(1, None)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"], "note": CPU_NOTE},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
class Env:
    """Process-group plumbing shared by every measurement of one bench run."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — pharmsol_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # keep stdout to the one JSON line without touching NCCL_DEBUG (the driver reads the communicator banner to
            # check the rank count): whatever NCCL prints while the communicator comes up goes to stderr
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                warm = torch.zeros(1, device=self.dev)
                dist.all_reduce(warm)
                torch.cuda.synchronize(self.dev)
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def maxr(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def minr(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN)

    def sumr(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)


def measure(env, name, nsub, nspp_total, steps, warmup, tol=1e-6, particles=1000, gather="auto", e2e=True, cpu_budget=0.0, force_nvrtc=False,
            peak_tf=None, clocks=False, pageable=True):
    """One workload through the resident path (+ e2e, + cpu baseline): returns a dict of measurements (rank 0 fills
    the host-side parts).  nspp_total columns are split over env.world ranks."""
    import pharmsol_b200 as ps
    from benches import harness as H
    from pharmsol_b200 import _lib
    torch = env.torch
    world, rank, dev = env.world, env.rank, env.dev
    cfg = WORKLOADS[name]
    w = make_workload(name, nsub, nspp_total, particles)
    compile_info = None
    if force_nvrtc:
        os.environ["PHARMSOL_B200_FORCE_NVRTC"] = "1"
    try:
        eq, data, ems = H.product_objects(w, device=env.local)
        if w["kind"] == "ode":
            eq.with_solver(getattr(ps.OdeSolver, cfg["solver"])).with_tolerances(tol, tol)
        if w["kind"] == "sde":
            eq.with_particles(particles).with_mode(ps.SdeMode.ParticleFilter).with_stepper(ps.EmMode.ReferenceAdaptive)
        t0 = time.perf_counter()
        source = eq._model.compile(eq._ctx())
        compile_info = {"module_source": source, "compile_ms": 1e3 * (time.perf_counter() - t0)}
    finally:
        if force_nvrtc:
            os.environ.pop("PHARMSOL_B200_FORCE_NVRTC", None)
    mode = default_gather(w["kind"], gather) if world > 1 else None
    job = ps.ResidentPsi(eq, data, w["support_points"], ems, device=dev, gather=(mode or "auto"))
    if world > 1:
        mode = job.gather_mode           # what was actually set up (a failed peer mapping falls back to NCCL)
    ctx = job.ctx
    npairs_rank = nsub * job.ncols
    npairs_total = nsub * nspp_total
    if peak_tf is None:
        peak_tf, _ = ctx.measure_fp64_peak()

    for _ in range(warmup):
        job.step()
    job.finish()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = ctx.launch_count
    kernel_ms_lib = []
    env.barrier()
    sampler = ClockSampler(env.local) if clocks else None
    if sampler:
        sampler.__enter__()
    t_wall0 = time.perf_counter()
    for k in range(steps):
        env.flush.zero_()                                   # L2 flush, outside the event bracket
        ev[k][0].record()
        job.step(after_compute=ev[k][1].record)             # launch(es) | event | gather (overlapped with the compute when pushed / phased)
        ev[k][2].record()
    env.barrier()
    t_wall = time.perf_counter() - t_wall0
    if sampler:
        sampler.__exit__()
    launches = ctx.launch_count - launches0
    psi = job.finish()
    kernel_ms_lib.append(ctx.last_kernel_ms)                # library-side events around the psi kernels of the last step
    step_ms = [e[0].elapsed_time(e[2]) for e in ev]
    kern_ms = [e[0].elapsed_time(e[1]) for e in ev]
    total_ms = env.maxr(float(np.sum(step_ms)))
    kernel_ms_avg = float(np.mean(kern_ms)) if mode != "push" else float(np.mean(kernel_ms_lib))
    kernel_ms_max = env.maxr(kernel_ms_avg)
    counters = ctx.last_counters                             # of the last launch (every launch does identical work)
    value = npairs_total * steps / (total_ms * 1e-3)
    n_nan = int(env.sumr(float(torch.isnan(psi).sum().item()))) // max(world, 1)      # every rank holds the whole matrix
    n_neginf = int(env.sumr(float(torch.isneginf(psi).sum().item()))) // max(world, 1)

    # ---- cross-rank check: recompute a block of ANOTHER rank's columns here and compare with what arrived ----------
    gather_ok = None
    if world > 1:
        other = (rank + 1) % world
        lo, hi = job.sharded.part.range(other)
        n = max(0, min(hi - lo, 128))
        same = 1.0
        if n > 0:
            spp_blk = torch.empty((job.nparams, n), dtype=torch.float64, device=dev)
            tmp = torch.full((n, nsub), float("nan"), dtype=torch.float64, device=dev)
            _lib.upload_support_points(ctx, w["support_points"][lo:lo + n], spp_blk.data_ptr(), n, job._stream())
            _lib.log_likelihood_matrix_device(ctx, eq._model, job.pop, spp_blk.data_ptr(), n, n, tmp.data_ptr(), nsub, lo, job._stream())
            torch.cuda.synchronize(dev)
            try:
                ctx.collect_errors()
            except ps.PharmsolError:
                pass
            a, b = tmp.view(torch.int64), psi.t()[lo:lo + n].contiguous().view(torch.int64)      # bit patterns: NaN == NaN
            same = 1.0 if bool(torch.equal(a, b)) else 0.0
        gather_ok = bool(env.minr(same) == 1.0)

    nobs = observations_per_subject(w)
    flops = algorithmic_flops(name, npairs_rank, counters, nobs)
    achieved_tf = flops / (kernel_ms_avg * 1e-3) * 1e-12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 8.0 * npairs_rank + 8.0 * job.nparams * job.ncols
    traffic = None
    try:    # measured DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), scaled per pair where the capture was a sample
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
        if tj:
            traffic = tj["per_pair"] * npairs_rank if "per_pair" in tj else (tj["bytes"] if (nsub, job.ncols) == (cfg["nsub"], cfg["nspp"]) else None)
    except Exception:
        traffic = None
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                "peak_source": "measured in this run: register-resident DFMA chains on all SMs (MEASURED_PEAKS.json carries no FP64 figure)",
                "kernel": "psi_entry_<model>_s<solver> (one thread per pair)", "kernel_ms": kernel_ms_avg,
                "algorithmic_flops_per_launch": flops, "flops_per_pair": flops / max(npairs_rank, 1),
                "device_counters": counters,
                "hbm": {"achieved": alg_bytes / (kernel_ms_avg * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (kernel_ms_avg * 1e-3) * 1e-9 / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    if name == "c2":
        ft = algorithmic_flops(name, npairs_rank, counters, nobs, true_count=True)
        roofline["frac_instruction_true"] = ft / (kernel_ms_avg * 1e-3) * 1e-12 / peak_tf
        roofline["note"] = ("frac follows SURVEY §8d's 2n(21+7+7)+60 flop per step; frac_instruction_true counts the b row once (FSAL: row 7 of A is b) "
                            "and 20 non-zero a-coefficients; ncu op counts (dfma/dadd/dmul) are in profiles/")
    if name == "c5":
        roofline["note"] = ("the SDE kernel is instruction-issue bound (Philox + Box-Muller + the step controller), not FP64: `frac` is the "
                            "ncu-counted FP64 work against the FP64 peak, `issue` the warp-instruction rate against 4 schedulers x SMs x clock")
        roofline["particle_evals_per_s"] = float(counters["evals"]) / (kernel_ms_avg * 1e-3)
        attempts = float(counters["steps"] + counters["rejected"])
        props = torch.cuda.get_device_properties(dev)
        clock_mhz = getattr(props, "clock_rate", 1965000) / 1e3           # the SM clock the bench's `clocks` samples confirm (1965 MHz)
        issue_peak = props.multi_processor_count * 4 * clock_mhz * 1e6
        issue_rate = attempts * C5_WARP_INST_PER_ATTEMPT / (kernel_ms_avg * 1e-3)
        roofline["issue"] = {"achieved": issue_rate, "peak": issue_peak, "unit": "warp-inst/s", "frac": issue_rate / issue_peak,
                             "warp_inst_per_attempt": C5_WARP_INST_PER_ATTEMPT, "source": "profiles/r02_final_c5_ncu.txt"}

    # ---- end to end through the public host-buffer API -------------------------------------------------
    e2e_rec, e2e_page = None, None
    if e2e:
        cols = job.columns                                  # this rank's columns
        spp_pinned, p1 = _lib.pinned_array((len(cols), job.nparams))
        spp_pinned[:] = w["support_points"][cols]
        out_pinned, p2 = _lib.pinned_array((nsub, len(cols)), order="F")
        nrep = max(1, min(steps, 5))
        for _ in range(0 if name == "c5" else 2):       # C5: 20 s per matrix; the resident steps above already warmed the kernel
            eq.log_likelihood_matrix(data, spp_pinned, ems, out=out_pinned)
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(nrep):
            eq.log_likelihood_matrix(data, spp_pinned, ems, out=out_pinned)
        torch.cuda.synchronize(dev)
        t_e2e = env.maxr(time.perf_counter() - t0)
        same = bool(np.array_equal(out_pinned, psi[:, torch.as_tensor(cols, device=psi.device)].cpu().numpy(), equal_nan=True))
        e2e_rec = {"value": npairs_total * nrep / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(spp_pinned.nbytes), "d2h_bytes_per_step": int(out_pinned.nbytes),
                   "ms_per_step": 1e3 * t_e2e / nrep, "api": "pharmsol_b200.log_likelihood_matrix (host numpy in / out, pinned) -> pharmsol_cuda_log_likelihood_matrix",
                   "matches_resident_result": same}
        if pageable and name != "c5":
            spp_page = np.array(spp_pinned)                  # malloc'ed, pageable: what a caller's Array2 is
            out_page = np.zeros((nsub, len(cols)), order="F")
            eq.log_likelihood_matrix(data, spp_page, ems, out=out_page)
            env.barrier()
            t0 = time.perf_counter()
            for _ in range(nrep):
                eq.log_likelihood_matrix(data, spp_page, ems, out=out_page)
            torch.cuda.synchronize(dev)
            t_pg = env.maxr(time.perf_counter() - t0)
            e2e_page = {"value": npairs_total * nrep / t_pg, "unit": UNIT, "ms_per_step": 1e3 * t_pg / nrep,
                        "buffers": "pageable (numpy / malloc) support points and psi", "matches_pinned_result": bool(np.array_equal(out_page, out_pinned, equal_nan=True))}
        _lib.host_free(p1); _lib.host_free(p2)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and cpu_budget > 0:
        try:
            r, info = cpu_reference_rate(name, w, nsub, tol, particles, cpu_budget)
            cpu = {"value": r, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"], "note": CPU_NOTE}
        except Exception as e:   # the baseline must never take the bench line down
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    rec = {"value": value, "unit": UNIT, "ms_per_step": total_ms / steps, "steps": steps, "warmup": warmup,
           "kernel_ms": kernel_ms_max, "gather_ms": max(0.0, total_ms / steps - kernel_ms_max), "gather_share": max(0.0, 1.0 - kernel_ms_max / (total_ms / steps)),
           "config": config_dict(name, w, nsub, job.ncols, nspp_total, world, tol, particles, gather_text(world, mode)),
           "e2e": e2e_rec, "e2e_pageable": e2e_page, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "psi_nan": n_nan, "psi_neg_inf": n_neginf, "wall_s_timed_region": t_wall, "compile": compile_info,
           "gather_matches_single_gpu": gather_ok}
    if sampler:
        rec["clocks"] = sampler.summary()
    rec["_objects"] = (w, eq, data, ems, psi, job)      # for the caller's follow-up measurements (not serialised)
    return rec


def slim(rec, keys=("value", "unit", "ms_per_step", "steps", "kernel_ms", "gather_ms", "gather_share", "config", "e2e", "e2e_pageable", "gpu_launches",
                    "roofline", "cpu_baseline", "psi_nan", "psi_neg_inf", "compile", "gather_matches_single_gpu")):
    out = {k: rec[k] for k in keys if k in rec}
    r = out.get("roofline")
    if r:       # keep the nested records compact: the full counters stay in the headline
        out["roofline"] = {k: r[k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel_ms", "flops_per_pair", "traffic", "particle_evals_per_s", "issue", "note") if k in r}
    return out


def library_multi_device(env, rec, steps):
    """rank 0: the headline matrix through ONE process and a multi-device context of the C ABI (the path a Rust caller
    of log_likelihood_matrix takes): host buffers in, each device's slab copied straight into the caller's matrix."""
    import pharmsol_b200 as ps
    from benches import harness as H
    from pharmsol_b200 import _lib
    w, eq, data, ems, psi, job = rec["_objects"]
    world = env.world
    out = None
    # The other ranks must leave their GPUs idle while rank 0 drives all of them from one process: an NCCL barrier would
    # keep a spinning kernel resident on every GPU (and time-slice against rank 0's launches there), so they block on the
    # CPU-side rendezvous store instead.
    store = env.dist.distributed_c10d._get_default_store()
    env.barrier()
    if env.rank != 0:
        store.wait(["pharmsol_b200_library_multi_device_done"])
    if env.rank == 0:
        try:
            eq2, data2, ems2 = H.product_objects(w, device=list(range(world)))
            if w["kind"] == "ode":
                eq2.with_solver(eq._solver).with_tolerances(eq._rtol, eq._atol)
            spp = w["support_points"]
            nsub = len(w["subjects"])
            spp_pinned, p1 = _lib.pinned_array(spp.shape)
            spp_pinned[:] = spp
            out_pinned, p2 = _lib.pinned_array((nsub, spp.shape[0]), order="F")
            for _ in range(2):
                eq2.log_likelihood_matrix(data2, spp_pinned, ems2, out=out_pinned)
            nrep = max(1, min(steps, 5))
            t0 = time.perf_counter()
            for _ in range(nrep):
                eq2.log_likelihood_matrix(data2, spp_pinned, ems2, out=out_pinned)
            dt = time.perf_counter() - t0
            same = bool(np.array_equal(out_pinned, psi.cpu().numpy(), equal_nan=True))
            ctx2 = eq2._ctx()
            pop2 = eq2.population(data2, ems2)
            _lib.log_likelihood_matrix_replicated(ctx2, eq2._model, pop2, spp_pinned)
            t0 = time.perf_counter()
            for _ in range(nrep):
                _lib.log_likelihood_matrix_replicated(ctx2, eq2._model, pop2, spp_pinned)
            dt_rep = time.perf_counter() - t0
            out = {"devices": world, "api": "pharmsol_cuda_ctx_create_multi + pharmsol_cuda_log_likelihood_matrix (one process, host buffers)",
                   "e2e_value": nsub * spp.shape[0] * nrep / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / nrep, "matches_torchrun_result": same,
                   "replicated_value": nsub * spp.shape[0] * nrep / dt_rep, "replicated_ms_per_step": 1e3 * dt_rep / nrep,
                   "replicated": "pharmsol_cuda_log_likelihood_matrix_replicated: psi resident on every device, chunks pushed by the copy engines"}
            _lib.host_free(p1); _lib.host_free(p2)
        except Exception as e:      # noqa: BLE001 - a failure here must not take the bench line down
            out = {"devices": world, "error": repr(e)}
        store.set("pharmsol_b200_library_multi_device_done", "1")
    env.barrier()
    return out


def run_product(args):
    env = Env()
    world, rank = env.world, env.rank
    cfg = WORKLOADS[args.workload]
    nsub = args.nsub or cfg["nsub"]
    ncols_arg = args.nspp or cfg["nspp"]
    nspp_total = ncols_arg if args.scaling == "strong" else ncols_arg * world
    steps, warmup = args.steps, max(args.warmup, 3)
    if args.workload == "c5":
        warmup = max(1, min(args.warmup, 1))

    import pharmsol_b200 as ps  # noqa: F401
    from pharmsol_b200 import _lib
    peak_tf, clk = _lib.context(env.local).measure_fp64_peak()

    head = measure(env, args.workload, nsub, nspp_total, steps, warmup, tol=args.tol, particles=args.particles, gather=args.gather, e2e=not args.no_e2e,
                   cpu_budget=0.0 if args.no_cpu_baseline else 15.0, force_nvrtc=args.force_nvrtc, peak_tf=peak_tf, clocks=True)
    ok = head["gather_matches_single_gpu"] in (None, True)

    extras = args.extras == "auto" and args.workload == "c2" and not args.nsub and not args.nspp and args.scaling == "weak"
    lib_multi = None
    configs, strong = None, None
    if world > 1 and extras:
        lib_multi = library_multi_device(env, head, steps)
    head.pop("_objects", None)
    if extras:
        env.torch.cuda.empty_cache()
        if world == 1:
            configs = {}
            for name, kw in (("c1", dict(steps=20, warmup=3)), ("c3", dict(steps=5, warmup=3)), ("c4", dict(steps=5, warmup=3, force_nvrtc=True)),
                             ("c5", dict(steps=1, warmup=1))):
                c = WORKLOADS[name]
                try:
                    r = measure(env, name, c["nsub"], c["nspp"], tol=args.tol, particles=args.particles, e2e=not args.no_e2e,
                                cpu_budget=0.0 if args.no_cpu_baseline else 4.0, peak_tf=peak_tf, **kw)
                    r.pop("_objects", None)
                    configs[name] = slim(r)
                except Exception as e:      # noqa: BLE001
                    configs[name] = {"error": repr(e)}
                env.torch.cuda.empty_cache()
        strong = {}
        for label, name, ns, total, kw in (("c2", "c2", 500, 20000, dict(steps=5, warmup=3)), ("c3_full", "c3", 10000, 50000, dict(steps=3, warmup=2))):
            try:
                r = measure(env, name, ns, total, tol=args.tol, particles=args.particles, gather=args.gather, e2e=False, peak_tf=peak_tf, **kw)
                r.pop("_objects", None)
                ok = ok and r["gather_matches_single_gpu"] in (None, True)
                s = slim(r, keys=("value", "unit", "ms_per_step", "steps", "kernel_ms", "gather_ms", "gather_share", "config", "gpu_launches", "psi_nan",
                                  "gather_matches_single_gpu"))
                s["scaling"] = "strong"
                s["roofline_frac"] = r["roofline"]["frac"]
                strong[label] = s
            except Exception as e:      # noqa: BLE001
                strong[label] = {"error": repr(e)}
            env.torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": head["config"], "clocks": head.get("clocks"), "e2e": head["e2e"], "e2e_pageable": head["e2e_pageable"],
                "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"],
                "psi_nan": head["psi_nan"], "psi_neg_inf": head["psi_neg_inf"], "wall_s_timed_region": head["wall_s_timed_region"],
                "fp64_peak_clock_mhz": clk, "kernel_ms": head["kernel_ms"], "gather_ms": head["gather_ms"], "gather_share": head["gather_share"],
                "compile": head["compile"], "gather_matches_single_gpu": head["gather_matches_single_gpu"]}
        if configs is not None:
            line["configs"] = configs
        if strong is not None:
            line["strong"] = strong
        if lib_multi is not None:
            line["library_multi_device"] = lib_multi
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()
    return 0 if ok else 3


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
