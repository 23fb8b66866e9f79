"""Bounds / undefined-behaviour check of the DEVICE SOURCE on the CPU: tests/hostsim built with AddressSanitizer + UBSan
(compute-sanitizer is closed on the GPU pool).  Walks every per-pair code path: closed forms through the timeline program
(staged copy, partial CTAs, warp-task index space), the lag cursor, all seven ODE solvers (BDF keeps its difference
array and change-of-step matrices in dynamically indexed local arrays), stiff model, SDE kernel.

    HOSTSIM_SANITIZE=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_odr_violation=0 python scripts/hostsim_sanitize.py
(every hostsim module carries its own copy of the shim's globals, hence detect_odr_violation=0)
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

assert os.environ.get("HOSTSIM_SANITIZE"), "set HOSTSIM_SANITIZE=1 (and LD_PRELOAD libasan)"
from benches import workloads as W
from hostsim import HostSim
import fixtures as FX


def ems(w):
    return [(1 if k == "additive" else 2, f, p) for _, (k, f, p) in w["error_models"].items()]


for name, nsub, nspp in (("c1", 4, 150), ("c1", 3, 40), ("c3", 3, 130)):
    w = W.make(name, nsub=nsub, nspp=nspp)
    psi, pred, info = HostSim(w["dsl"]).set_subjects(w["subjects"]).run(w["support_points"], ems(w), want_pred=True)
    assert info["code"] == 0 and np.all(np.isfinite(psi)), (name, info)
    print(name, nsub, nspp, "ok", flush=True)
w = W.make("c2", nsub=3, nspp=40)
hs = HostSim(w["dsl"]).set_subjects(w["subjects"])
for solver in ("Dopri5", "Tsit45", "Sdirk4", "TrBdf2", "Rodas4", "Bdf", "Esdirk34"):
    for tol in (1e-4, 1e-8):
        psi, _, info = hs.run(w["support_points"], ems(w), solver=solver, rtol=tol, atol=tol)
        assert info["code"] == 0 and np.all(np.isfinite(psi)), (solver, info)
    print(solver, "ok", flush=True)
w = W.make("c4", nsub=3, nspp=40)
hs = HostSim(w["dsl"]).set_subjects(w["subjects"])
for solver in ("Rodas4", "Bdf", "Esdirk34", "Sdirk4"):
    psi, _, info = hs.run(w["support_points"], ems(w), solver=solver, rtol=1e-6, atol=1e-6)
    assert info["code"] == 0, (solver, info)
    print("c4", solver, "ok", flush=True)
src, twin, p, ops, _ = FX.CORPUS["analytical_full"]
_, pred, info = HostSim(src).set_subjects([("s", ops)]).run(np.array([p]), None, want_pred=True)
assert info["code"] == 0
for kernel in ("two_compartments_with_absorption", "three_compartments_cl"):
    rng = np.random.default_rng(5)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_hostsim import _random_subject
    subjects = [(f"r{i}", _random_subject(rng, kernel.endswith("with_absorption"), 3)) for i in range(6)]
    spp = np.stack([rng.uniform(0.05, 2.0, 20) if n not in ("v", "vc", "vp", "v2", "v3") else rng.uniform(5, 80, 20) for n in FX.KERNEL_PARAMS[kernel]], axis=1)
    _, _, info = HostSim(FX.kernel_dsl(kernel)).set_subjects(subjects).run(spp, [(1, 0.05, (0.1, 0.15, 0, 0))], want_pred=True)
    assert info["code"] in (0, 12), info
    print(kernel, "ok", flush=True)
w = W.make("c5", nsub=2, nspp=3, particles=48)
hs = HostSim(w["dsl"]).set_subjects(w["subjects"])
for mode in (0, 1):
    for em in (0, 1):
        psi, _, info = hs.run(w["support_points"], ems(w), particles=48, sde_mode=mode, em_mode=em, em_dt=0.05, sde_normals=mode)
        assert not np.any(np.isnan(psi))
print("HOSTSIM_SANITIZE PASS")
