"""Tuning harness (run under gpurun): times one workload's psi kernel for a list of compile / launch
variants through the NVRTC path.  Each variant = (extra NVRTC flags, block size).

    python scripts/tune.py c2 --nspp 20000 --variants "PSI_MIN_BLOCKS=4|128" "PSI_MIN_BLOCKS=6|128"
"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--nsub", type=int, default=0)
ap.add_argument("--nspp", type=int, default=0)
ap.add_argument("--tol", type=float, default=1e-6)
ap.add_argument("--solver", default="")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--sde-fp64", action="store_true", help="SDE: FP64 Box-Muller noise instead of the FP32 default")
ap.add_argument("--particles", type=int, default=1000)
ap.add_argument("--variants", nargs="+", default=["|128"])
args = ap.parse_args()

import bench
import pharmsol_b200 as ps
from benches import harness as H, workloads as W

cfg = bench.WORKLOADS[args.workload]
kw = dict(nsub=args.nsub or cfg["nsub"], nspp=args.nspp or cfg["nspp"])
w = W.make(args.workload, **kw)
dev = torch.device("cuda", 0)
os.environ["PHARMSOL_B200_FORCE_NVRTC"] = "1"
ref = None
for v in args.variants:
    flags, block = v.split("|")
    os.environ["PHARMSOL_B200_NVRTC_FLAGS"] = " ".join("-D" + t for t in flags.split())   # "NAME=VAL NAME2=VAL2|block"
    os.environ["PHARMSOL_B200_BLOCK"] = block or "128"
    os.environ["PHARMSOL_B200_CUBIN_CACHE"] = tempfile.mkdtemp()
    eq, data, ems = H.product_objects(w, device=0)
    if w["kind"] == "ode":
        eq.with_solver(getattr(ps.OdeSolver, args.solver or cfg["solver"])).with_tolerances(args.tol, args.tol)
    if w["kind"] == "sde":
        eq.with_particles(args.particles).with_mode(ps.SdeMode.ParticleFilter).with_stepper(ps.EmMode.ReferenceAdaptive).with_noise_precision(args.sde_fp64)
    job = ps.ResidentPsi(eq, data, w["support_points"], ems, device=dev, shard=False)
    for _ in range(2):
        job.launch()
    psi = job.finish().clone()
    ms = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); job.launch(); e1.record()
        job.finish()
        ms.append(e0.elapsed_time(e1))
    if ref is None:
        ref = psi
    dmax = float((psi - ref).abs().max().item())
    c = job.ctx.last_counters
    print(json.dumps({"variant": v, "ms": min(ms), "pairs_per_s": kw["nsub"] * kw["nspp"] / (min(ms) * 1e-3), "steps": c["steps"], "rejected": c["rejected"],
                      "evals": c["evals"], "max_abs_diff_vs_first": dmax}), flush=True)
