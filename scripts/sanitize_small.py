"""A small tour of every kernel path, sized for compute-sanitizer (run under gpurun):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
closed forms through the staged timeline program (full and partial CTAs, warp-task and diagonal index spaces), the lag
path, all seven ODE solvers, SDE (both steppers, particle filter, batch), the two-shard entry points and the [0, 0]
multi-device context, predictions in chunks."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np

import pharmsol_b200 as ps
from benches import harness as H, workloads as W
from pharmsol_b200 import _lib
import fixtures as FX

os.environ["PHARMSOL_B200_PRED_CHUNK_KB"] = "64"
ok = True
for name, nsub, nspp in (("c1", 5, 150), ("c1", 3, 40), ("c3", 4, 200)):
    w = W.make(name, nsub=nsub, nspp=nspp)
    eq, data, ems = H.product_objects(w)
    a = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    p, _ = eq.predictions_matrix(data, w["support_points"])
    eq2, data2, ems2 = H.product_objects(w, device=[0, 0])
    b = ps.log_likelihood_matrix(eq2, data2, w["support_points"], ems2)
    ok = ok and np.array_equal(a, b) and np.all(np.isfinite(a)) and np.all(np.isfinite(p))
    print(name, nsub, nspp, "ok", ok, flush=True)
w = W.make("c2", nsub=4, nspp=160)
for solver in ("Dopri5", "Tsit45", "Sdirk4", "TrBdf2", "Rodas4", "Bdf", "Esdirk34"):
    eq, data, ems = H.product_objects(w)
    eq.with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(1e-6, 1e-6)
    a = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    ok = ok and np.all(np.isfinite(a))
    print(solver, "ok", ok, flush=True)
out = ps.log_likelihood_batch(eq, data, w["support_points"][:4], ps.ResidualErrorModels().add(0, ps.ResidualErrorModel.combined(0.1, 0.2)))
ok = ok and np.all(np.isfinite(out))
src, twin, p, ops, _ = FX.CORPUS["analytical_full"]          # lag + fa: the per-thread event cursor
pred = ps.Equation.from_dsl(src).estimate_predictions(ps.Subject("s", ops), p).flat_predictions()
ok = ok and np.all(np.isfinite(pred))
w = W.make("c5", nsub=3, nspp=6, particles=64)
eq, data, ems = H.product_objects(w)
for mode in (ps.SdeMode.MeanPrediction, ps.SdeMode.ParticleFilter):
    for em, dt in ((ps.EmMode.FixedStep, 0.05), (ps.EmMode.ReferenceAdaptive, 0.05)):
        eq.with_particles(64).with_mode(mode).with_stepper(em, dt)
        a = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
        ok = ok and not np.any(np.isnan(a))
eq.with_noise_precision(True)
a = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
out = ps.log_likelihood_batch(eq, data, w["support_points"][:3], ps.ResidualErrorModels().add(0, ps.ResidualErrorModel.constant(0.5)))
ok = ok and not np.any(np.isnan(a)) and out.shape == (3,)
print("SANITIZE_SMALL", "PASS" if ok else "FAIL", flush=True)
sys.exit(0 if ok else 1)
