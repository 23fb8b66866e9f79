"""SURVEY §8 f.2 as written: the host artifact with the reference's FROZEN compiled-backend symbols
(src/dsl/compiled_backend_abi.rs:6-33; loader src/dsl/aot.rs:316-353, 404-470).  The DSL emitter writes the host twin
of the device model; g++ builds the cdylib; these tests `dlopen` it the way `load_aot_model` does (version check,
model-info envelope, one 7-argument `extern "C"` function per role) and check every role's arithmetic.  CPU only."""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CORPUS = json.load(open(os.path.join(ROOT, "tests", "golden", "dsl_corpus.json")))
FROZEN = ["pharmsol_dsl_api_version", "pharmsol_dsl_model_info_json_ptr", "pharmsol_dsl_model_info_json_len"]
ROLES = ["derive", "dynamics", "outputs", "init", "drift", "diffusion", "route_lag", "route_bioavailability"]

FEATURE = """
name = feature
kind = ode
params = ka, ke, v, tlag, f
covariates = wt
states = depot, central
derived = cl, scaled
outputs = cp, amount

bolus(po) -> depot
infusion(iv) -> central
lag(po) = tlag * (wt / 70)
fa(po) = f

cl = ke * v * (wt / 70)^0.75
scaled = central / v

init(depot) = 0
init(central) = 5 * v

dx(depot) = -ka * depot
dx(central) = ka * depot - (cl / v) * central

out(cp) = scaled ~ continuous()
out(amount) = central + depot ~ continuous()
"""


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_frozen_symbols_version_and_envelope(ps, tmp_path):
    path = ps.compile_module_source_to_native_aot(FEATURE, tmp_path / "feature.pkm")
    syms = exported(path)
    for s in FROZEN + ["pharmsol_dsl_kernel_" + r for r in ("derive", "dynamics", "outputs", "init", "route_lag", "route_bioavailability")]:
        assert s in syms, s
    assert "pharmsol_dsl_kernel_drift" not in syms and "pharmsol_dsl_kernel_diffusion" not in syms
    lib = C.CDLL(str(path))
    lib.pharmsol_dsl_api_version.restype = C.c_uint32
    assert lib.pharmsol_dsl_api_version() == 2                       # AOT_API_VERSION, aot.rs:43
    art = ps.NativeArtifact(path)
    env = art.envelope                                               # CompiledModelInfoEnvelope, compiled_backend_abi.rs:81-86
    assert set(env) == {"abi_version", "model", "functions"} and env["abi_version"] == 2
    assert env["functions"] == {"derive": True, "dynamics": True, "outputs": True, "init": True, "drift": False, "diffusion": False,
                                "route_lag": True, "route_bioavailability": True}
    m = env["model"]                                                 # NativeModelInfo, model_info.rs:17-92
    assert m["name"] == "feature" and m["kind"] == "Ode" and m["parameters"] == ["ka", "ke", "v", "tlag", "f"]
    assert m["state_len"] == 2 and m["derived_len"] == 2 and m["output_len"] == 2 and m["route_len"] >= 1
    assert [r["name"] for r in m["routes"]] == ["po", "iv"] and m["routes"][0]["has_lag"] and m["routes"][0]["has_bioavailability"]
    assert m["routes"][1]["inject_input_to_destination"] is True


def test_every_role_computes_the_model(ps, tmp_path):
    art = ps.NativeArtifact(ps.compile_module_source_to_native_aot(FEATURE, tmp_path / "feature.pkm"))
    ka, ke, v, tlag, f, wt = 1.3, 0.21, 42.0, 0.75, 0.8, 81.0
    p, cov, x = [ka, ke, v, tlag, f], [wt], [12.0, 30.0]
    cl = ke * v * (wt / 70.0) ** 0.75
    d = art.call("derive", 1.0, x, p, cov, [0.0, 0.0], [0.0, 0.0], out_len=2)
    assert d[0] == pytest.approx(cl, rel=1e-15) and d[1] == pytest.approx(30.0 / v, rel=1e-15)
    # derive may write in place: `out` aliases `derived` (compiled_backend_abi.rs:147-180)
    buf = np.zeros(2)
    art.call("derive", 1.0, x, p, cov, [0.0, 0.0], buf, out=buf)
    assert np.array_equal(buf, d)
    # dynamics does NOT add the infusion rate: the reference runtime injects route inputs itself
    routes = [0.0, 7.0]
    dx = art.call("dynamics", 1.0, x, p, cov, routes, d, out_len=2)
    assert dx[0] == pytest.approx(-ka * 12.0, rel=1e-15) and dx[1] == pytest.approx(ka * 12.0 - (cl / v) * 30.0, rel=1e-14)
    y = art.call("outputs", 1.0, x, p, cov, routes, d, out_len=2)
    assert y[0] == pytest.approx(30.0 / v, rel=1e-15) and y[1] == 42.0
    # init writes the states in place: `out` aliases `states`
    st = np.zeros(2)
    art.call("init", 0.0, st, p, cov, [0.0, 0.0], d, out=st)
    assert st[0] == 0.0 and st[1] == 5 * v
    # route properties: only the declaring route's slot is written; the caller pre-fills 0 / 1 (native.rs:941-1018)
    m = art.info
    slot = m["routes"][0]["index"]
    lag = art.call("route_lag", 2.0, [0, 0], p, cov, [0.0] * m["route_len"], d, out=np.full(m["route_len"], -7.0))
    assert lag[slot] == pytest.approx(tlag * wt / 70.0, rel=1e-15)
    fa = art.call("route_bioavailability", 2.0, [0, 0], p, cov, [0.0] * m["route_len"], d, out=np.full(m["route_len"], 1.0))
    assert fa[slot] == f
    other = [k for k in range(m["route_len"]) if k != slot]
    assert all(lag[k] == -7.0 and fa[k] == 1.0 for k in other)


def test_sde_and_analytical_models_export_their_roles(ps, tmp_path):
    from benches import workloads as W
    sde = ps.NativeArtifact(ps.compile_module_source_to_native_aot(W.model_source("c5_one_cpt_sde"), tmp_path / "sde.pkm"))
    assert sde.envelope["functions"]["drift"] and sde.envelope["functions"]["diffusion"] and not sde.envelope["functions"]["dynamics"]
    assert sde.info["kind"] == "Sde" and {"drift", "diffusion", "outputs"} <= set(sde.functions)
    names = sde.info["parameters"]
    p = [0.5 + 0.1 * k for k in range(len(names))]
    dx = sde.call("drift", 0.0, [10.0], p, routes=[0.0] * max(sde.info["route_len"], 1), out_len=1)
    assert dx[0] == pytest.approx(-p[names.index("ke")] * 10.0, rel=1e-15)
    ana = ps.NativeArtifact(ps.compile_module_source_to_native_aot(W.model_source("c3_three_cpt_cov"), tmp_path / "ana.pkm"))
    assert ana.info["kind"] == "Analytical" and ana.info["analytical"] == "ThreeCompartmentsWithAbsorption"
    assert ana.envelope["functions"]["derive"] and not ana.envelope["functions"]["dynamics"] and "dynamics" not in ana.functions


def test_twin_dynamics_reproduce_the_closed_form_through_rk4(ps, tmp_path):
    """The exported `dynamics` drives an RK4 loop here (what the reference runtime does with diffsol) and must land on
    the analytic two-compartment-with-absorption solution the oracle / device are checked against."""
    from benches import workloads as W
    art = ps.NativeArtifact(ps.compile_module_source_to_native_aot(W.model_source("c2_two_cpt_oral_ode"), tmp_path / "c2.pkm"))
    ka, ke, kcp, kpc, v = 1.1, 0.15, 0.08, 0.05, 40.0
    p = [ka, ke, kcp, kpc, v]
    f = lambda t, x: art.call("dynamics", t, x, p, routes=[0.0] * max(art.info["route_len"], 1), out_len=3).copy()
    x, h = np.array([100.0, 0.0, 0.0]), 1.0 / 256
    for k in range(int(6.0 / h)):
        t = k * h
        k1 = f(t, x); k2 = f(t + h / 2, x + h / 2 * k1); k3 = f(t + h / 2, x + h / 2 * k2); k4 = f(t + h, x + h * k3)
        x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    from scipy.linalg import expm
    A = np.array([[-ka, 0, 0], [ka, -(ke + kcp), kpc], [0, kcp, -kpc]])
    want = expm(A * 6.0) @ np.array([100.0, 0.0, 0.0])
    assert np.max(np.abs(x - want) / np.abs(want)) <= 1e-9
    y = art.call("outputs", 6.0, x, p, out_len=1)
    assert y[0] == pytest.approx(want[1] / v, rel=1e-9)


ACCEPTED = [r for r in CORPUS if r["expect"] == "accept"]


@pytest.mark.parametrize("row", ACCEPTED[::3], ids=lambda r: r["where"])
def test_corpus_models_build_as_native_artifacts(ps, tmp_path, row):
    """Every third DSL source the reference accepts: the host twin compiles, exports exactly the roles the envelope
    declares, and `outputs` (the one required symbol) runs on zero buffers without touching anything out of range."""
    eq = ps.Equation.from_dsl(row["source"])
    path = eq._model.export_host_artifact(tmp_path / "m.pkm")
    syms = exported(path)
    art = ps.NativeArtifact(path)
    for role in ROLES:
        assert art.envelope["functions"][role] == (("pharmsol_dsl_kernel_" + role) in syms), role
    m = art.info
    n = lambda k: max(int(m[k]), 1)
    out = art.call("outputs", 0.0, [0.0] * n("state_len"), [1.0] * max(len(m["parameters"]), 1), [1.0] * max(len(m["covariates"]), 1),
                   [0.0] * n("route_len"), [0.0] * n("derived_len"), out_len=n("output_len"))
    assert out.shape == (n("output_len"),)
