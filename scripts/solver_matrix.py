"""Every ODE solver on the two ODE BASELINE workloads at full size (run under gpurun): ms per matrix, failures, and the
worst scaled log-likelihood difference against the RODAS4 / Dopri5 result at the same tolerance."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pharmsol_b200 as ps
from benches import harness as H, workloads as W

dev = torch.device("cuda", 0)
for name, nobs in (("c2", 12), ("c4", 8)):
    cfg = bench.WORKLOADS[name]
    w = W.make(name, nsub=cfg["nsub"], nspp=cfg["nspp"])
    ref = None
    for solver in ("Dopri5", "Tsit45", "Rodas4", "Sdirk4", "TrBdf2", "Esdirk34", "Bdf"):
        eq, data, ems = H.product_objects(w, device=0)
        eq.with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(1e-6, 1e-6)
        job = ps.ResidentPsi(eq, data, w["support_points"], ems, device=dev, shard=False)
        rec = {"workload": name, "solver": solver, "rtol": 1e-6}
        try:
            job.launch(); job.finish()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); job.launch(); e1.record()
            psi = job.finish()
            rec["ms"] = e0.elapsed_time(e1)
            rec["pairs_per_s"] = cfg["nsub"] * cfg["nspp"] / (rec["ms"] * 1e-3)
            rec["nan"] = int(torch.isnan(psi).sum().item())
            c = job.ctx.last_counters
            rec["steps_per_pair"] = c["steps"] / (cfg["nsub"] * cfg["nspp"])
            if ref is None:
                ref = psi.clone()
            rec["max_scaled_dll_vs_first"] = float(((psi - ref).abs() / (ref.abs() + nobs)).max().item())
        except ps.PharmsolError as e:
            rec["error"] = str(e)
        print(json.dumps(rec), flush=True)
