"""First end-to-end GPU check: small versions of the five configs vs the oracle (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import pharmsol_b200 as ps
from pharmsol_b200 import _lib
from benches import workloads as W, harness as H

ctx = _lib.context(0)
print("fp64 peak TFLOP/s, clock:", ctx.measure_fp64_peak())

def run(name, nsub, nspp, solver=None, tol=None, oracle_kw=None, **kw):
    w = W.make(name, nsub=nsub, nspp=nspp, **kw)
    eq, data, ems = H.product_objects(w)
    if solver is not None:
        eq.with_solver(solver).with_tolerances(tol, tol)
    t0 = time.time()
    try:
        psi = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    except ps.PharmsolError as e:
        print(name, "GPU ERROR", e, e.pair); return
    t1 = time.time()
    psi2 = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    t2 = time.time()
    print(f"{name}: first call {t1-t0:.3f}s second {t2-t1:.4f}s kernel {ctx.last_kernel_ms:.3f} ms counters {ctx.last_counters}")
    om, od, oe = H.oracle_objects(w, **(oracle_kw or {}))
    ref, info = om.log_likelihood_matrix(od, w["support_points"], oe, return_info=True)
    err = H.rel_err(psi, ref, 1e-12)
    print(f"   oracle {info['seconds']:.3f}s  max rel err {err.max():.3e}  median {np.median(err):.3e}  psi[0,:3]={psi[0,:3]} ref={ref[0,:3]}")
    pred, offs = eq.predictions_matrix(data, w["support_points"][:4])
    op = np.array([om.predictions(od.subjects[0], w["support_points"][j]) for j in range(4)]).T
    n0 = offs[1]
    print("   pred rel err subj0:", H.rel_err(pred[:n0], op, 1e-12).max())

run("c1", 64, 256)
run("c3", 32, 128)
run("c2", 16, 128, solver=ps.OdeSolver.Dopri5, tol=1e-8, oracle_kw=dict(solver="dopri5", rtol=1e-10, atol=1e-10))
run("c2", 16, 128, solver=ps.OdeSolver.Tsit45, tol=1e-8, oracle_kw=dict(solver="tsit45", rtol=1e-10, atol=1e-10))
run("c4", 8, 64, solver=ps.OdeSolver.Sdirk4, tol=1e-8, oracle_kw=dict(solver="dopri5", rtol=1e-10, atol=1e-10))
run("c4", 8, 64, solver=ps.OdeSolver.TrBdf2, tol=1e-6, oracle_kw=dict(solver="dopri5", rtol=1e-10, atol=1e-10))
run("c4", 8, 64, solver=ps.OdeSolver.Dopri5, tol=1e-8, oracle_kw=dict(solver="dopri5", rtol=1e-10, atol=1e-10))
# full-size timing
for name, nsub, nspp, kw in [("c1", 1000, 1000, {}), ("c2", 500, 20000, dict(solver=ps.OdeSolver.Dopri5, tol=1e-6)), ("c3", 1000, 5000, {})]:
    w = W.make(name, nsub=nsub, nspp=nspp)
    eq, data, ems = H.product_objects(w)
    if "solver" in kw: eq.with_solver(kw["solver"]).with_tolerances(kw["tol"], kw["tol"])
    for _ in range(3):
        t0 = time.time(); psi = ps.log_likelihood_matrix(eq, data, w["support_points"], ems); t1 = time.time()
        print(f"{name} {nsub}x{nspp}: e2e {t1-t0:.4f}s kernel {ctx.last_kernel_ms:.3f} ms -> {nsub*nspp/ctx.last_kernel_ms*1e3:.3e} pairs/s; counters {ctx.last_counters}; finite {np.isfinite(psi).all()}")
