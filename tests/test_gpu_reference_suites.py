"""The reference's remaining integration suites, restated against the CUDA path through the same public calls
(`estimate_predictions`, solver selection, runtime backends):

* tests/numerical_stability.rs:49-136  — analytical kernels vs their ODE twins (reference bar: abs 1e-2 or rel 1e-2)
* tests/test_solvers.rs:67-103         — every OdeSolver variant vs Bdf (reference bar: abs 1e-2)
* tests/full_feature_dsl_backend_parity.rs:136-183 — runtime-jit vs runtime-native-aot on the corpus: same model
  info, same predictions, both equal to the handwritten twin (here: NVRTC vs the `.pkm` artifact vs the oracle twin)
"""
import math

import numpy as np
import pytest

import fixtures as FX

pytestmark = pytest.mark.gpu


def _preds(ps, eq, ops, params):
    return np.array(eq.estimate_predictions(ps.Subject("s", ops), params).flat_predictions())


def _missing(times):
    return [("missing_observation", float(t), "cp") for t in times]


T13 = (0.0, 1.0, 2.0, 4.0, 8.0, 12.0, 24.0, 25.0, 26.0, 27.0, 28.0, 32.0, 36.0)
STABILITY = {
    # numerical_stability.rs:138-151, 153-232
    "infusion": dict(
        ops=[("bolus", 0.0, 100.0, "load"), ("infusion", 24.0, 150.0, "iv", 3.0)] + _missing(T13), p=[0.1, 1.0],
        ana="name = infusion_reference\nkind = analytical\nparams = ke, v\nstates = central\noutputs = cp\nbolus(load) -> central\ninfusion(iv) -> central\n"
            "structure = one_compartment\nout(cp) = central / v ~ continuous()\n",
        ode="name = infusion_reference_ode\nkind = ode\nparams = ke, v\nstates = central\noutputs = cp\nbolus(load) -> central\ninfusion(iv) -> central\n"
            "dx(central) = -ke * central\nout(cp) = central / v ~ continuous()\n"),
    # numerical_stability.rs:234-314
    "absorption": dict(
        ops=[("bolus", 0.0, 100.0, "oral"), ("infusion", 24.0, 150.0, "iv", 3.0), ("bolus", 48.0, 100.0, "load")]
            + _missing(T13 + (48.0, 49.0, 50.0, 52.0, 56.0, 60.0)), p=[1.0, 0.1, 1.0],
        ana="name = absorption_reference\nkind = analytical\nparams = ka, ke, v\nstates = gut, central\noutputs = cp\nbolus(load) -> central\nbolus(oral) -> gut\n"
            "infusion(iv) -> central\nstructure = one_compartment_with_absorption\nout(cp) = central / v ~ continuous()\n",
        ode="name = absorption_reference_ode\nkind = ode\nparams = ka, ke, v\nstates = gut, central\noutputs = cp\nbolus(load) -> central\nbolus(oral) -> gut\n"
            "infusion(iv) -> central\ndx(gut) = -ka * gut\ndx(central) = ka * gut - ke * central\nout(cp) = central / v ~ continuous()\n"),
    # numerical_stability.rs:316-382
    "two_compartment": dict(
        ops=[("bolus", 0.0, 100.0, "load"), ("infusion", 24.0, 150.0, "iv", 3.0)] + _missing(T13), p=[0.1, 3.0, 1.0, 1.0],
        ana="name = two_comp_reference\nkind = analytical\nparams = ke, kcp, kpc, v\nstates = central, peripheral\noutputs = cp\nbolus(load) -> central\n"
            "infusion(iv) -> central\nstructure = two_compartments\nout(cp) = central / v ~ continuous()\n",
        ode="name = two_comp_reference_ode\nkind = ode\nparams = ke, kcp, kpc, v\nstates = central, peripheral\noutputs = cp\nbolus(load) -> central\n"
            "infusion(iv) -> central\ndx(central) = -ke * central - kcp * central + kpc * peripheral\ndx(peripheral) = kcp * central - kpc * peripheral\n"
            "out(cp) = central / v ~ continuous()\n"),
}


@pytest.mark.parametrize("label", sorted(STABILITY))
def test_numerical_stability_analytical_vs_ode(ps, label):
    c = STABILITY[label]
    expected = _preds(ps, ps.Equation.from_dsl(c["ana"]), c["ops"], c["p"])
    actual = _preds(ps, ps.Equation.from_dsl(c["ode"]), c["ops"], c["p"])             # reference defaults: rtol = atol = 1e-4
    assert expected.shape == actual.shape and np.all(np.isfinite(expected))
    abs_err = np.abs(expected - actual)
    rel_err = abs_err / np.maximum(np.abs(expected), 1e-2)
    assert np.all((abs_err <= 1e-2) | (rel_err <= 1e-2))                              # the reference's own bar
    tight = _preds(ps, ps.Equation.from_dsl(c["ode"]).with_tolerances(1e-10, 1e-10), c["ops"], c["p"])
    assert np.max(np.abs(tight - expected) / np.maximum(np.abs(expected), 1e-2)) <= 1e-8


def test_infusion_stability_case_closed_form(ps):
    """The `infusion` case has an elementary solution; pins both model families, not just their agreement."""
    c = STABILITY["infusion"]
    ke, v = c["p"]

    def cp(t):
        if t == 0.0:
            return 0.0          # the observation sorts before the same-time bolus (data/event.rs:292-304)
        x = 100.0 * math.exp(-ke * t)
        if t > 24.0:
            on = min(t, 27.0) - 24.0
            x += 50.0 / ke * (1.0 - math.exp(-ke * on)) * math.exp(-ke * (t - 24.0 - on))
        return x / v
    want = np.array([cp(t) for t in T13])
    assert np.allclose(_preds(ps, ps.Equation.from_dsl(c["ana"]), c["ops"], c["p"]), want, rtol=1e-12, atol=1e-12)
    assert np.allclose(_preds(ps, ps.Equation.from_dsl(c["ode"]).with_tolerances(1e-11, 1e-11), c["ops"], c["p"]), want, rtol=1e-8, atol=1e-8)


# ---- tests/test_solvers.rs ------------------------------------------------------------------------------------------
SOLVER_OPS = [("bolus", 0.0, 100.0, "iv_bolus"), ("infusion", 12.0, 200.0, "iv", 2.0)] + \
             [("observation", t, 0.0, "cp") for t in (0.5, 2.0, 8.0, 12.5, 14.0, 24.0)]
SOLVER_SRC = ("name = solver_selection_one_cpt\nkind = ode\nparams = ke, v\nstates = central\noutputs = cp\nbolus(iv_bolus) -> central\ninfusion(iv) -> central\n"
              "dx(central) = -ke * central\nout(cp) = central / v ~ continuous()\n")


def _solver_preds(ps, solver):
    return _preds(ps, ps.Equation.from_dsl(SOLVER_SRC).with_solver(solver), SOLVER_OPS, [0.1, 50.0])


def test_bdf_produces_predictions(ps):
    p = _solver_preds(ps, ps.OdeSolver.Bdf)
    assert len(p) == 6 and np.all(np.isfinite(p))


@pytest.mark.parametrize("name", ["Tsit45", "TrBdf2", "Esdirk34", "Dopri5", "Sdirk4", "Rodas4"])
def test_solver_matches_bdf(ps, name):
    ref_p = _solver_preds(ps, ps.OdeSolver.Bdf)
    test_p = _solver_preds(ps, getattr(ps.OdeSolver, name))
    assert ref_p.shape == test_p.shape
    assert np.max(np.abs(ref_p - test_p)) < 0.01          # test_solvers.rs:72-103


# ---- tests/full_feature_dsl_backend_parity.rs ---------------------------------------------------------------------
@pytest.mark.parametrize("case", ["ode_full", "analytical_full", "ode", "analytical", "sde"])
def test_corpus_backends_agree(ps, oracle, case, tmp_path):
    from pharmsol_b200 import _lib
    src, twin, p, ops, _ = FX.CORPUS[case]
    jit = ps.compile_module_source_to_runtime(src, ps.RuntimeCompilationTarget.Jit)
    aot = ps.compile_module_source_to_runtime(src, ps.RuntimeCompilationTarget.CudaAot(tmp_path / f"{case}.pkm"))
    assert aot._model.compile(_lib.context(0)) == "artifact" and jit._model.compile(_lib.context(0)) != "artifact"
    assert jit.info == aot.info == ps.read_aot_model_info(tmp_path / f"{case}.pkm")["model"]
    a, b = _preds(ps, jit, ops, p), _preds(ps, aot, ops, p)
    assert np.array_equal(a, b)
    if case != "sde":
        kw = dict(solver="dopri5", rtol=1e-12, atol=1e-12) if case.startswith("ode") else {}
        want = oracle.Model(twin, **kw).predictions(oracle.Subject(ops), p)
        bar = 1e-4 if case.startswith("ode") else 1e-8      # runtime_corpus.rs:186-196
        assert np.max(np.abs(a - want) / np.maximum(np.abs(want), 1e-8)) <= bar


# ---- tests/bimodal_ke_entrypoint_matrix.rs + tests/support/bimodal_ke.rs ------------------------------------------------
BIMODAL_KE = ("name = bimodal_ke\nkind = ode\n\nparams = ke, v\nstates = central\noutputs = cp\n\ninfusion(iv) -> central\n\n"
              "dx(central) = -ke * central\n\nout(cp) = central / v ~ continuous()\n")


def test_bimodal_ke_entrypoint_matrix(ps, tmp_path):
    """Three ways to the same runtime model (bimodal_ke.rs:278-330): runtime Jit, runtime AOT round trip, direct AOT
    export + load_runtime_artifact; each must reproduce the reference predictions (here the closed form)."""
    times = (0.5, 1.0, 2.0, 3.0, 4.0, 6.0, 8.0)
    ke, v = 1.2, 50.0
    ops = [("infusion", 0.0, 500.0, "iv", 0.5)] + [("missing_observation", t, "cp") for t in times]
    want = np.array([1000.0 / ke * (1.0 - math.exp(-ke * 0.5)) * math.exp(-ke * (t - 0.5)) / v for t in times])
    tight = lambda e: e.with_tolerances(1e-11, 1e-11)      # noqa: E731
    jit = tight(ps.compile_module_source_to_runtime(BIMODAL_KE, ps.RuntimeCompilationTarget.Jit))
    assert jit.backend() == ps.RuntimeBackend.Jit
    path = ps.compile_module_source_to_aot(BIMODAL_KE, tmp_path / "bimodal-ke-direct-aot.pkm", configure=tight)
    direct = ps.load_runtime_artifact(path, ps.RuntimeArtifactFormat.CudaAot)
    runtime_aot = tight(ps.compile_module_source_to_runtime(BIMODAL_KE, ps.RuntimeCompilationTarget.CudaAot(tmp_path / "bimodal-ke-runtime-aot.pkm")))
    assert direct.backend() == runtime_aot.backend() == ps.RuntimeBackend.CudaAot
    got = [_preds(ps, m, ops, [ke, v]) for m in (jit, direct, runtime_aot)]
    for g in got:
        assert np.max(np.abs(g - want)) <= 1e-9
    assert np.array_equal(got[0], got[1]) and np.array_equal(got[0], got[2])
    # numeric data labels reach the same model only through the input_<n> / outeq_<n> aliases, never by position
    with pytest.raises(ps.PharmsolError):
        _preds(ps, jit, [("infusion", 0.0, 500.0, "0", 0.5), ("missing_observation", 1.0, "cp")], [ke, v])


def test_full_feature_public_shape(ps):
    """full_feature_dsl_backend_parity.rs:26-134"""
    ode = ps.Equation.from_dsl(FX.ODE_FULL_SOURCE).info
    assert ode["name"] == "ode_full_feature_parity"
    assert ode["parameters"] == ["ka", "ke", "kcp", "kpc", "v", "tlag", "f_oral", "base_depot", "base_central", "base_peripheral"]
    assert [c["name"] for c in ode["covariates"]] == ["wt", "renal"]
    assert [r["name"] for r in ode["routes"]] == ["oral", "load", "iv"]
    assert [r["declaration_index"] for r in ode["routes"]] == [0, 1, 2]
    assert [r["index"] for r in ode["routes"]] == [0, 1, 0]
    assert [o["name"] if isinstance(o, dict) else o for o in ode["outputs"]] == ["cp"]
    ana = ps.Equation.from_dsl(FX.ANALYTICAL_FULL_SOURCE).info
    assert ana["name"] == "analytical_full_feature_parity"
    assert ana["parameters"] == ["ka", "ke", "v", "tlag", "f_oral", "base_gut", "base_central"]
    assert ana["derived"] == ["adjusted_v"]
    assert [r["index"] for r in ana["routes"]] == [0, 1, 0]
