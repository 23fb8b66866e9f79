#!/usr/bin/env python
"""bench.py — psi evaluations/sec (subject x support-point pairs) of the B200 psi-matrix backend.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...    # the path's CPU implementation on host cores

A "step" is one full psi matrix: every (subject, support point) pair of the workload simulated and
its log-likelihood stored (column-major), with the population, the support points and the output
resident in HBM.  Default workload = BASELINE.json configs[1]: two-compartment oral absorption
`ode!` (Dopri5), 500 subjects x 20,000 support points, 10 doses + 12 observations per subject.
For N > 1 (torchrun, one rank per GPU) every rank owns 20,000 support-point columns (weak scaling)
and the step ends with the in-place NCCL all-gather of the psi column slabs.

Timing: W >= 3 untimed warm-up steps; K timed steps, each bracketed by CUDA events on the launch
stream; an L2 flush (a 512 MiB device memset, outside the event bracket) separates iterations;
barrier + synchronize on both sides of the timed region; max over ranks.  `e2e` times the same
step through the public host-buffer API (H2D of the support points + D2H of psi inside the region).
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
import threading
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
ERK_STEP = lambda n: 2.0 * n * (21 + 7 + 7) + 60.0          # Dopri5 / Tsit5 stage combinations + error norm + controller


def algorithmic_flops(name, npairs, counters, nobs_per_subject, nsteps_per_subject):
    """Algorithmic FP64 work of one psi matrix (flop-equivalents, SURVEY §8d convention)."""
    if name == "c1":      # counters["evals"] = closed-form propagation steps executed on the device
        return counters["evals"] * (24 + 5) + npairs * (nobs_per_subject * 5 + 20)
    if name == "c3":
        return counters["evals"] * (850 + 75 + 4) + npairs * nobs_per_subject * 5
    if name == "c2":
        attempts = counters["steps"] + counters["rejected"]
        return counters["evals"] * F_RHS["c2"] + attempts * ERK_STEP(3) + npairs * nobs_per_subject * 15
    if name == "c4":
        # SDIRK: per Newton iteration one RHS + one 2x2 solve; per step one Jacobian + LU
        return counters["evals"] * F_RHS["c4"] + counters["newton"] * 30 + (counters["steps"] + counters["rejected"]) * 150
    if name == "c5":
        return float(counters["evals"]) * 60.0
    raise KeyError(name)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--nsub", type=int, default=0)
    ap.add_argument("--nspp", type=int, default=0, help="support points PER GPU")
    ap.add_argument("--tol", type=float, default=1e-6, help="ODE rtol = atol")
    ap.add_argument("--particles", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: fused all-gather by peer stores from the psi kernel, a separate NCCL all-gather, or auto "
                         "(fused for ODE / SDE models, NCCL for closed-form models whose psi is produced at GB/s rates)")
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


def make_workload(args, world):
    from benches import workloads as W
    cfg = WORKLOADS[args.workload]
    nsub = args.nsub or cfg["nsub"]
    nspp_per_gpu = args.nspp or cfg["nspp"]
    kw = dict(nsub=nsub, nspp=nspp_per_gpu * world)
    if args.workload == "c5":
        kw["particles"] = args.particles
    w = W.make(args.workload, **kw)
    return w, nsub, nspp_per_gpu


def config_dict(args, w, nsub, nspp_per_gpu, world):
    cfg = {"workload": f"{args.workload}: {w['desc']}", "nsub": nsub, "nspp_per_gpu": nspp_per_gpu, "nspp_total": nspp_per_gpu * world,
           "sharding": f"support-point columns x{world}",
           "l2": "512 MiB device memset between timed iterations (outside the event bracket)"}
    if w["kind"] == "ode":
        cfg.update(solver=WORKLOADS[args.workload]["solver"], rtol=args.tol, atol=args.tol)
    if w["kind"] == "sde":
        cfg.update(particles=args.particles, sde_mode="particle filter (SDE::estimate_log_likelihood, sde/mod.rs:526-577)",
                   stepper="reference adaptive Euler-Maruyama (sde/em.rs:134-167)")
    return cfg


def events_per_subject(w):
    ops = w["subjects"][0][1]
    nobs = sum(1 for o in ops if o[0] == "observation")
    nev = sum(1 for o in ops if o[0] in ("observation", "bolus", "infusion"))
    ninf = sum(1 for o in ops if o[0] == "infusion")
    return nobs, (nev - 1) + 2 * ninf       # propagation steps: event intervals + infusion boundary splits


# ---------------------------------------------------------------------------------------------------
def host_cores():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1: ask the OS, not OpenMP)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(args, w, nsub, budget_s, threads=0):
    """Time the oracle (restated CPU path, OpenMP over subjects) on a bounded sample: all subjects x
    the first S support points, S calibrated so the sample takes ~budget_s.  Returns (pairs/s, info)."""
    from benches import harness as H
    import oracle as O
    threads = threads or host_cores()
    kw = {}
    if w["kind"] == "ode":
        kw = dict(solver="dopri5", rtol=args.tol, atol=args.tol)
    if w["kind"] == "sde":
        kw = dict(particles=args.particles)
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    w, nsub, nspp_per_gpu = make_workload(args, 1)
    total = args.steps + args.warmup
    budget = min(15.0, max(1.0, 150.0 / max(total, 1)))
    for _ in range(args.warmup):
        cpu_reference_rate(args, w, nsub, budget)
    rates, info = [], None
    t_all = 0.0
    pairs = 0
    for _ in range(args.steps):
        r, info = cpu_reference_rate(args, w, nsub, budget)
        rates.append(r)
        t_all += info["seconds"]
        pairs += nsub * info["nspp_sample"]
    value = pairs / t_all
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, w, nsub, nspp_per_gpu, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"],
                             "note": "restated CPU oracle (C++/OpenMP over subjects, the reference's rayon decomposition); the Rust reference cannot be built here (no cargo)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
def run_product(args):
    import torch
    import torch.distributed as dist
    import pharmsol_b200 as ps
    from benches import harness as H

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — pharmsol_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: the image exports NCCL_DEBUG=VERSION, which prints a banner to stdout
        os.environ["NCCL_DEBUG"] = os.environ.get("PHARMSOL_B200_NCCL_DEBUG", "WARN")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)                      # anything NCCL prints while the communicator comes up goes to stderr
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w, nsub, nspp_per_gpu = make_workload(args, world)
    eq, data, ems = H.product_objects(w, device=local)
    cfg = WORKLOADS[args.workload]
    if w["kind"] == "ode":
        eq.with_solver(getattr(ps.OdeSolver, cfg["solver"])).with_tolerances(args.tol, args.tol)
    if w["kind"] == "sde":
        eq.with_particles(args.particles).with_mode(ps.SdeMode.ParticleFilter).with_stepper(ps.EmMode.ReferenceAdaptive)
    job = ps.ResidentPsi(eq, data, w["support_points"], ems, device=dev, peer_stores=("auto" if args.gather == "auto" else args.gather == "peer"))
    fused = getattr(job.sharded, "peer_ptrs", None) is not None
    ctx = job.ctx
    npairs_rank = nsub * job.ncols
    npairs_total = nsub * nspp_per_gpu * world
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    peak_tf, clk = ctx.measure_fp64_peak()

    # ---- resident-in-HBM timing -------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        job.step()
    job.finish()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ctx.launch_count
    barrier()
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            flush.zero_()                                   # L2 flush, outside the event bracket
            ev[k][0].record()
            job.step(after_compute=ev[k][1].record)         # launch(es) | event | all-gather (overlapped with the tail phase when phased)
            ev[k][2].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - launches0
    psi = job.finish()
    step_ms = [e[0].elapsed_time(e[2]) for e in ev]
    kern_ms = [e[0].elapsed_time(e[1]) for e in ev]
    total_ms = maxr(float(np.sum(step_ms)))
    kernel_ms_avg = float(np.mean(kern_ms))
    counters = ctx.last_counters                             # of the last launch (every launch does identical work)
    value = npairs_total * args.steps / (total_ms * 1e-3)
    n_nan = int(torch.isnan(psi).sum().item())               # NaN marks a failed pair (none expected)
    n_neginf = int(torch.isneginf(psi).sum().item())         # -inf is a legitimate particle-filter result (sde/mod.rs:699-703)

    nobs, nsteps = events_per_subject(w)
    flops = algorithmic_flops(args.workload, npairs_rank, counters, nobs, nsteps)
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
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tj:
            traffic = tj["per_pair"] * npairs_rank if "per_pair" in tj else (tj["bytes"] if (nsub, nspp_per_gpu) == (cfg["nsub"], cfg["nspp"]) else None)
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

    # ---- end to end through the public host-buffer API -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        from pharmsol_b200 import _lib
        cols = job.columns                                  # this rank's columns (two ranges when the gather is phased)
        spp_pinned, p1 = _lib.pinned_array((len(cols), job.nparams))
        spp_pinned[:] = w["support_points"][cols]
        out_pinned, p2 = _lib.pinned_array((nsub, len(cols)), order="F")
        for _ in range(2):
            eq.log_likelihood_matrix(data, spp_pinned, ems, out=out_pinned)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eq.log_likelihood_matrix(data, spp_pinned, ems, out=out_pinned)
        torch.cuda.synchronize(dev)
        t_e2e = maxr(time.perf_counter() - t0)
        same = bool(np.array_equal(out_pinned, psi[:, torch.as_tensor(cols, device=psi.device)].cpu().numpy()))
        e2e = {"value": npairs_total * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(spp_pinned.nbytes), "d2h_bytes_per_step": int(out_pinned.nbytes),
               "ms_per_step": 1e3 * t_e2e / args.steps, "api": "pharmsol_b200.log_likelihood_matrix (host numpy in / out, pinned) -> pharmsol_cuda_log_likelihood_matrix",
               "matches_resident_result": same}
        _lib.host_free(p1); _lib.host_free(p2)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r, info = cpu_reference_rate(args, w, nsub, 15.0)
            cpu = {"value": r, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"],
                   "note": "restated CPU oracle (C++/OpenMP over subjects); the Rust reference cannot be built here (no cargo)"}
        except Exception as e:   # the baseline must never take the bench line down
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": dict(config_dict(args, w, nsub, nspp_per_gpu, world),
                                                    gather=("none (1 GPU)" if world == 1 else "fused: psi kernel stores to every rank over NVLink + device barrier"
                                                            if fused else "NCCL all_gather_into_tensor (in place)" + (", 7/8 of the columns gathered while the last 1/8 computes" if len(job.ranges) > 1 else ""))),
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "psi_nan": n_nan, "psi_neg_inf": n_neginf, "wall_s_timed_region": t_wall, "fp64_peak_clock_mhz": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
