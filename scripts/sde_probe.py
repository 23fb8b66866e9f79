"""SDE kernel probe (run under gpurun): time per pair for both steppers and both likelihood modes."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import pharmsol_b200 as ps
from benches import harness as H, workloads as W
nsub, nspp, npart = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w = W.make("c5", nsub=nsub, nspp=nspp, particles=npart)
for em, dt, mode in [(ps.EmMode.FixedStep, 0.1, ps.SdeMode.ParticleFilter), (ps.EmMode.FixedStep, 0.1, ps.SdeMode.MeanPrediction),
                     (ps.EmMode.FixedStep, 0.02, ps.SdeMode.ParticleFilter), (ps.EmMode.ReferenceAdaptive, 0.1, ps.SdeMode.ParticleFilter),
                     (ps.EmMode.ReferenceAdaptive, 0.1, ps.SdeMode.MeanPrediction)]:
    eq, data, ems = H.product_objects(w, device=0)
    eq.with_particles(npart).with_mode(mode).with_stepper(em, dt)
    job = ps.ResidentPsi(eq, data, w["support_points"], ems, shard=False)
    job.launch(); job.finish()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); job.launch(); e1.record(); psi = job.finish()
    ms = e0.elapsed_time(e1); c = job.ctx.last_counters
    print(json.dumps(dict(em=em, dt=dt, mode=mode, ms=ms, pairs_per_s=nsub * nspp / ms * 1e3, particle_steps_per_s=c["evals"] / ms * 1e3, counters=c,
                          finite=bool(torch.isfinite(psi).all()), mean_ll=float(psi.mean()))), flush=True)
