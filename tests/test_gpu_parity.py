"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI (ctypes ->
libpharmsol_cuda.so), against the CPU oracle and the committed goldens on the same inputs.

Tolerances (north_star): analytical <= 1e-12 relative; ODE <= 1e-6 relative on predictions and
log-likelihood (solver tolerances tightened accordingly); SDE statistical (|d mean ll| <= 4 SE).
Log-likelihood sums can cancel to ~0, so their relative error is measured against
|ll| + n_obs (each observation contributes O(1) terms): `ll_close`.
"""
import math
import os

import numpy as np
import pytest

import fixtures as FX
import golden_math
from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from benches import harness
    return harness


@pytest.fixture(scope="module")
def W():
    from benches import workloads
    return workloads


def rel(a, b, floor):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def ll_close(gpu, ref, nobs, tol):
    gpu, ref = np.asarray(gpu, float), np.asarray(ref, float)
    err = np.abs(gpu - ref) / (np.abs(ref) + nobs)
    assert np.all(np.isfinite(gpu)), "non-finite psi"
    assert err.max() <= tol, f"max scaled ll error {err.max():.3e} > {tol:g} at {np.unravel_index(err.argmax(), err.shape)}"
    return err.max()


def gpu_predictions(ps, eq, ops, params):
    """estimate_predictions for one subject -> 1-D array (missing observations included)."""
    return np.array(eq.estimate_predictions(ps.Subject("s", ops), params).flat_predictions())


def test_native_library_is_loaded_and_device_is_sm100(ps):
    from pharmsol_b200 import _lib
    assert ps.device_count() >= 1
    ctx = _lib.context(0)
    before = ctx.launch_count
    tf, clk = ctx.measure_fp64_peak()
    assert tf > 10.0 and clk > 1000.0
    assert ctx.launch_count > before
    maps = open("/proc/self/maps").read()
    assert "libpharmsol_cuda.so" in maps


# ---- analytical kernels on the reference's unit-test fixtures ---------------------------------------
@pytest.mark.parametrize("kernel", list(FX.KERNEL_PARAMS))
def test_builtin_kernel_fixture(ps, oracle, kernel):
    """equation/analytical/*_models.rs fixtures: DSL `structure = <kernel>` on the GPU vs the oracle's
    literal kernel (and the expm golden for the six rate-constant kernels)."""
    base = kernel.replace("_cl", "")
    params, ops = FX.KERNEL_FIXTURES[base]
    if "_cl" in kernel:   # same dynamics expressed as clearances (volumes 1 except where the fixture says otherwise)
        params = {"one_compartment_cl": [0.1, 1.0], "one_compartment_cl_with_absorption": [1.0, 0.1, 1.0],
                  "two_compartments_cl": [0.1, 3.0, 1.0, 3.0], "two_compartments_cl_with_absorption": [1.0, 0.1, 3.0, 1.0, 3.0],
                  "three_compartments_cl": [0.1, 3.0, 2.0, 1.0, 3.0, 4.0],
                  "three_compartments_cl_with_absorption": [1.0, 0.1, 3.0, 2.0, 1.0, 3.0, 4.0]}[kernel]
    eq = ps.Equation.from_dsl(FX.kernel_dsl(kernel))
    got = gpu_predictions(ps, eq, ops, params)
    want = oracle.Model(kernel).predictions(oracle.Subject(ops), params)
    assert got.shape == want.shape
    assert rel(got, want, 1e-10).max() <= 1e-12
    if "_cl" not in kernel:
        g = next(t for t in golden("timelines") if t["kernel"] == kernel)
        assert np.allclose(got, g["predictions"], rtol=1e-12, atol=1e-13)


def test_seq_of_kernels_random_steps_vs_expm_golden(ps):
    """kernels.json: one propagation step per case = bolus of x0 at t=0 (+ infusion) then one
    observation per state is not expressible through outputs for all states, so check the central
    compartment: x_central(dt) / 1."""
    cases = [c for c in golden("kernels") if c["rate"] == 0.0]
    by_kernel = {}
    for c in cases:
        by_kernel.setdefault(c["kernel"], []).append(c)
    for kernel, cs in by_kernel.items():
        absorb = kernel.endswith("with_absorption")
        eq = ps.Equation.from_dsl(FX.kernel_dsl(kernel))
        central = 1 if absorb else 0
        for c in cs:
            x = c["x"]
            # only the first one/two states can be loaded by boluses (inputs 0 and 1)
            loadable = 2 if absorb else 1
            if any(abs(v) > 0 for v in x[loadable:]):
                x = list(x[:loadable]) + [0.0] * (len(x) - loadable)
            ops = [("bolus", 0.0, x[0], "0")] + ([("bolus", 0.0, x[1], "1")] if absorb else []) + [("missing_observation", c["dt"], "0")]
            got = gpu_predictions(ps, eq, ops, c["p"] + [1.0])[0]
            # independent truth through scipy expm on the same (possibly truncated) initial state
            want = golden_math.step(kernel, c["p"], np.array(x), c["dt"], 0.0)[central]
            assert got == pytest.approx(want, rel=1e-11, abs=1e-12 * max(1.0, max(abs(v) for v in x)))


# ---- the reference's runtime corpus (DSL vs handwritten twin) ------------------------------------------
@pytest.mark.parametrize("case", ["analytical", "analytical_full"])
def test_corpus_analytical(ps, oracle, case):
    src, twin, p, ops, _ = FX.CORPUS[case]
    got = gpu_predictions(ps, ps.Equation.from_dsl(src), ops, p)
    want = oracle.Model(twin).predictions(oracle.Subject(ops), p)
    assert rel(got, want, 1e-10).max() <= 1e-12   # reference's own bar: 1e-8 (runtime_corpus.rs:186-196)


@pytest.mark.parametrize("case", ["ode", "ode_full"])
@pytest.mark.parametrize("solver", ["Dopri5", "Tsit45", "Sdirk4", "TrBdf2", "Rodas4", "Bdf", "Esdirk34"])
def test_corpus_ode(ps, oracle, case, solver):
    src, twin, p, ops, _ = FX.CORPUS[case]
    tol = 1e-8 if solver == "TrBdf2" else 1e-10
    eq = ps.Equation.from_dsl(src).with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(tol, tol)
    got = gpu_predictions(ps, eq, ops, p)
    want = oracle.Model(twin, solver="dopri5", rtol=1e-12, atol=1e-12).predictions(oracle.Subject(ops), p)
    assert rel(got, want, 1e-8).max() <= (1e-5 if solver in ("TrBdf2", "Bdf") else 1e-6)   # reference's own bar: 1e-4


def test_corpus_sde_zero_diffusion_is_deterministic(ps, oracle):
    src, twin, p, ops, _ = FX.CORPUS["sde"]
    got = gpu_predictions(ps, ps.Equation.from_dsl(src), ops, p)
    want = oracle.Model(twin).predictions(oracle.Subject(ops), p)
    assert np.all(np.isfinite(got))
    assert rel(got, want, 1e-8).max() <= 1e-6     # reference's own bar: 1e-4


# ---- the five BASELINE configs at oracle-checkable sizes ----------------------------------------------------
def _matrix(ps, H, w, **eqkw):
    eq, data, ems = H.product_objects(w)
    for k, v in eqkw.items():
        getattr(eq, k)(*v)
    return eq, data, ems, ps.log_likelihood_matrix(eq, data, w["support_points"], ems)


def test_c1_matrix(ps, oracle, H, W):
    w = W.make("c1", nsub=48, nspp=300)
    eq, data, ems, psi = _matrix(ps, H, w)
    assert psi.shape == (48, 300) and psi.flags.f_contiguous
    om, od, oe = H.oracle_objects(w)
    ref = om.log_likelihood_matrix(od, w["support_points"], oe)
    ll_close(psi, ref, 10, 1e-12)
    pred, offs = eq.predictions_matrix(data, w["support_points"][:9])
    assert pred.shape == (48 * 10, 9) and offs[-1] == 480
    for i in (0, 17, 47):
        for j in (0, 8):
            want = om.predictions(od.subjects[i], w["support_points"][j])
            assert rel(pred[offs[i]:offs[i + 1], j], want, 1e-12).max() <= 1e-12


@pytest.mark.parametrize("mode", ["interval_end", "interval_length"])
def test_c3_matrix_both_covariate_time_semantics(ps, oracle, H, W, mode):
    w = W.make("c3", nsub=24, nspp=160)
    w["oracle_model"] = "c3_three_cpt_cov_" + mode
    eq, data, ems = H.product_objects(w)
    eq.with_cov_time(ps.CovTime.IntervalEnd if mode == "interval_end" else ps.CovTime.IntervalLength)
    psi = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    om, od, oe = H.oracle_objects(w)
    ref = om.log_likelihood_matrix(od, w["support_points"], oe)
    # 3-cpt roots (atan2/cos/sin/pow) + ~20 chained steps: SURVEY §7 budget ~1e-12 on predictions;
    # the likelihood amplifies prediction error by |z| * pred/sigma, hence 1e-10 on the scaled ll
    ll_close(psi, ref, 10, 1e-10)
    pred, offs = eq.predictions_matrix(data, w["support_points"][:6])
    worst = 0.0
    for i in range(0, 24, 5):
        for j in range(6):
            want = om.predictions(od.subjects[i], w["support_points"][j])
            worst = max(worst, rel(pred[offs[i]:offs[i + 1], j], want, 1e-9).max())
    assert worst <= 1e-11, worst


@pytest.mark.parametrize("solver", ["Dopri5", "Tsit45"])
def test_c2_matrix_vs_closed_form(ps, oracle, H, W, solver):
    w = W.make("c2", nsub=12, nspp=160)
    eq, data, ems, psi = _matrix(ps, H, w, with_solver=(getattr(ps.OdeSolver, solver),), with_tolerances=(1e-10, 1e-10))
    # truth: the closed-form two_compartments_with_absorption kernel (the reference's own ODE<->analytical test)
    w2 = dict(w, oracle_model=w["oracle_truth_model"])
    om, od, oe = H.oracle_objects(w2)
    ref = om.log_likelihood_matrix(od, w["support_points"], oe)
    ll_close(psi, ref, 12, 1e-6)
    pred, offs = eq.predictions_matrix(data, w["support_points"][:8])
    for i in (0, 5, 11):
        for j in range(8):
            want = om.predictions(od.subjects[i], w["support_points"][j])
            assert rel(pred[offs[i]:offs[i + 1], j], want, 1e-9).max() <= 1e-6


def test_c2_reference_default_tolerance_agrees_with_oracle_solver(ps, oracle, H, W):
    """rtol = atol = 1e-4 (ode/mod.rs:40-41): both sides are ~1e-4 solvers, compare loosely."""
    w = W.make("c2", nsub=6, nspp=96)
    eq, data, ems, psi = _matrix(ps, H, w, with_solver=(ps.OdeSolver.Dopri5,), with_tolerances=(1e-4, 1e-4))
    om, od, oe = H.oracle_objects(dict(w, oracle_model=w["oracle_truth_model"]))
    ref = om.log_likelihood_matrix(od, w["support_points"], oe)
    pred, offs = eq.predictions_matrix(data, w["support_points"][:8])
    want = np.array([om.predictions(od.subjects[0], w["support_points"][j]) for j in range(8)]).T
    assert rel(pred[:offs[1]], want, 1e-3).max() <= 5e-3


@pytest.mark.parametrize("solver,tol,bar", [("Sdirk4", 1e-9, 1e-6), ("TrBdf2", 1e-8, 2e-5), ("Dopri5", 1e-10, 1e-6), ("Rodas4", 1e-9, 1e-6),
                                            ("Bdf", 1e-10, 1e-5), ("Esdirk34", 1e-10, 1e-6)])
def test_c4_stiff_vs_radau_golden(ps, solver, tol, bar):
    """stiff_c4.json: SciPy Radau rtol=1e-12 predictions (ke0 up to 50 /h)."""
    from benches import workloads
    eq = ps.Equation.from_dsl(workloads.model_source("c4_mm_effect")).with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(tol, tol)
    for c in golden("stiff_c4"):
        got = gpu_predictions(ps, eq, [tuple(o) for o in c["ops"]], c["params"])
        assert rel(got, c["predictions"], 1e-6).max() <= bar, (solver, c["params"])


@pytest.mark.parametrize("solver", ["Sdirk4", "Rodas4"])
def test_c4_matrix(ps, oracle, H, W, solver):
    w = W.make("c4", nsub=8, nspp=96)
    eq, data, ems, psi = _matrix(ps, H, w, with_solver=(getattr(ps.OdeSolver, solver),), with_tolerances=(1e-9, 1e-9))
    om, od, oe = H.oracle_objects(w, solver="dopri5", rtol=1e-11, atol=1e-12)
    ref = om.log_likelihood_matrix(od, w["support_points"], oe)
    ll_close(psi, ref, 8, 1e-6)


def test_c5_sde_mean_prediction_and_particle_filter_statistics(ps, oracle, H, W):
    """SDE parity is statistical: the reference RNG is an unseeded thread-local ChaCha (parity unpinned
    at rand 0.10), so compare seed-averaged log-likelihoods: |mean_gpu - mean_oracle| <= 4 SE."""
    w = W.make("c5", nsub=2, nspp=6, particles=512)
    eq, data, ems = H.product_objects(w)
    eq.with_particles(512)
    om, od, oe = H.oracle_objects(w, particles=512)
    nseed = 24
    for mode, omode in ((ps.SdeMode.MeanPrediction, 0), (ps.SdeMode.ParticleFilter, 1)):
        eq.with_mode(mode)
        g = np.stack([ps.log_likelihood_matrix(eq.with_seed(1000 + s), data, w["support_points"], ems) for s in range(nseed)])
        o = np.stack([om.log_likelihood_matrix(od, w["support_points"], oe, seed=77 + s, sde_mode=omode) for s in range(nseed)])
        assert np.all(np.isfinite(g))
        se = np.sqrt(g.var(axis=0, ddof=1) / nseed + o.var(axis=0, ddof=1) / nseed)
        z = np.abs(g.mean(axis=0) - o.mean(axis=0)) / np.maximum(se, 1e-9 * (1 + np.abs(o.mean(axis=0))))
        assert z.max() <= 4.5, (mode, z.max())
        # same seed -> same stream -> bit-identical result (Philox counters are pure functions of the pair)
        a = ps.log_likelihood_matrix(eq.with_seed(5), data, w["support_points"], ems)
        b = ps.log_likelihood_matrix(eq.with_seed(5), data, w["support_points"], ems)
        assert np.array_equal(a, b)


# ---- likelihood features: censoring, missing, per-observation error polynomials, occasions ---------------------
def test_censoring_missing_errorpoly_occasions(ps, oracle, W):
    dsl = W.model_source("c1_one_cpt_iv")
    subjects = []
    rng = np.random.default_rng(7)
    for i in range(6):
        ops = [("infusion", 0.0, 400.0 + 20 * i, "iv", 0.5), ("observation", 0.5, 3.1 + 0.1 * i, "cp"),
               ("missing_observation", 1.0, "cp"),
               ("censored_observation", 2.0, 0.5, "cp", "bloq"), ("censored_observation", 3.0, 6.0, "cp", "aloq"),
               ("observation_with_error", 4.0, 1.2, "cp", (0.05, 0.2, 0.0, 0.01), "none"),
               ("observation_with_error", 6.0, 0.3, "cp", (0.02, 0.1, 0.0, 0.0), "bloq"),
               ("reset",),
               ("infusion", 0.0, 300.0, "iv", 1.0), ("observation", 1.0, 2.0 + rng.random(), "cp"), ("observation", 5.0, 0.9, "cp")]
        subjects.append((f"s{i}", ops))
    spp = np.stack([np.linspace(0.1, 1.2, 40), np.linspace(40, 250, 40)], axis=1)
    for kind, em in (("additive", ("additive", 0.3, (0.1, 0.1, 0.0, 0.0))), ("proportional", ("proportional", 1.5, (0.05, 0.1, 0.01, 0.0)))):
        eq = ps.Equation.from_dsl(dsl)
        data = ps.Data([ps.Subject(i, o) for i, o in subjects])
        model = (ps.AssayErrorModel.additive if kind == "additive" else ps.AssayErrorModel.proportional)(ps.ErrorPoly(*em[2]), em[1])
        psi = ps.log_likelihood_matrix(eq, data, spp, ps.AssayErrorModels().add("cp", model))
        om = oracle.Model("one_cpt_iv")
        od = oracle.Data([oracle.Subject(o, i) for i, o in subjects])
        ref = om.log_likelihood_matrix(od, spp, oracle.ErrorModels([em]))
        ll_close(psi, ref, 8, 1e-12)


def test_extreme_censoring_takes_asymptotic_branch(ps, oracle, W):
    """|z| > 37 (distributions.rs:60-70, 95-103): finite through the asymptotic branch on both sides."""
    ops = [("infusion", 0.0, 500.0, "iv", 0.5), ("censored_observation", 1.0, 0.01, "cp", "bloq"), ("censored_observation", 2.0, 50.0, "cp", "aloq")]
    spp = np.array([[0.3, 100.0], [0.05, 30.0]])
    eq = ps.Equation.from_dsl(W.model_source("c1_one_cpt_iv"))
    em = ("additive", 0.0, (0.01, 0.0, 0.0, 0.0))
    psi = ps.log_likelihood_matrix(eq, ps.Data([ps.Subject("a", ops)]), spp,
                                   ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(*em[2]), 0.0)))
    ref = oracle.Model("one_cpt_iv").log_likelihood_matrix(oracle.Data([oracle.Subject(ops)]), spp, oracle.ErrorModels([em]))
    assert np.all(np.isfinite(ref)) and ref.min() < -1000
    ll_close(psi, ref, 2, 1e-12)


def test_lag_reorders_events_per_support_point(ps, oracle):
    """structs.rs:611-690: lag shifts bolus times per support point (threads of one warp take
    different event orders); fa scales at the lagged time."""
    src, twin, p, _, _ = FX.CORPUS["analytical"]
    ops = [("bolus", 0.0, 100.0, "oral"), ("bolus", 1.0, 50.0, "oral"), ("bolus", 6.0, 80.0, "oral")] + \
          [("observation", t, 1.0 + 0.1 * t, "cp") for t in (0.5, 1.0, 1.5, 2.0, 3.0, 6.0, 6.5, 8.0, 12.0)]
    lags = np.linspace(0.0, 2.5, 64)
    spp = np.stack([np.full(64, 1.0), np.full(64, 0.15), np.full(64, 25.0), lags, np.linspace(0.5, 1.0, 64)], axis=1)
    eq = ps.Equation.from_dsl(src)
    data = ps.Data([ps.Subject("a", ops)])
    em = ("additive", 0.0, (0.1, 0.1, 0.0, 0.0))
    psi = ps.log_likelihood_matrix(eq, data, spp, ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(*em[2]), 0.0)))
    ref = oracle.Model(twin).log_likelihood_matrix(oracle.Data([oracle.Subject(ops)]), spp, oracle.ErrorModels([em]))
    ll_close(psi, ref, 9, 1e-12)
    pred, _ = eq.predictions_matrix(data, spp)
    om = oracle.Model(twin)
    want = np.array([om.predictions(oracle.Subject(ops), s) for s in spp]).T
    assert rel(pred, want, 1e-10).max() <= 1e-12


@pytest.mark.parametrize("solver", ["Dopri5", "Tsit45", "Sdirk4", "TrBdf2", "Rodas4", "Bdf", "Esdirk34"])
def test_ode_infusion_dose_conservation(ps, solver):
    """ode/mod.rs:1274-1344."""
    src = "name = acc\nkind = ode\nparams = ke, v\nstates = central\noutputs = cp\ninfusion(iv) -> central\ndx(central) = -ke * central\nout(cp) = central / v ~ continuous()\n"
    eq = ps.Equation.from_dsl(src).with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(1e-6, 1e-6)
    p = [0.0, 1.0]
    P = lambda ops: gpu_predictions(ps, eq, ops, p)
    assert P([("infusion", 0.0, 100.0, "iv", 0.1), ("observation", 0.5, 0.0, "cp")])[0] == pytest.approx(100.0, rel=1e-4)
    pr = P([("infusion", 0.0, 100.0, "iv", 0.1), ("observation", 0.1, 0.0, "cp"), ("observation", 0.5, 0.0, "cp")])
    assert pr[0] == pytest.approx(100.0, rel=1e-4) and pr[1] == pytest.approx(100.0, rel=1e-4)
    assert P([("infusion", 0.0, 100.0, "iv", 0.01), ("observation", 0.01, 0.0, "cp")])[0] == pytest.approx(100.0, rel=1e-4)
    pr = P([("observation", 0.0, 0.0, "cp"), ("infusion", 0.5, 100.0, "iv", 0.01), ("observation", 0.52, 0.0, "cp")])
    assert pr[0] == 0.0 and pr[1] == pytest.approx(100.0, rel=1e-4)
    pr = P([("infusion", 0.0, 100.0, "iv", 0.5), ("infusion", 0.5, 100.0, "iv", 0.5), ("observation", 1.0, 0.0, "cp")])
    assert pr[0] == pytest.approx(200.0, rel=1e-4)


def test_likelihood_case_of_the_reference(ps, oracle):
    # tests/ode_optimizations.rs:1105-1184
    c = FX.LIKELIHOOD_CASE
    em = ps.AssayErrorModels().add("outeq_0", ps.AssayErrorModel.additive(ps.ErrorPoly(*c["error_model"][2]), 0.0))
    eq = ps.Equation.from_dsl(FX.kernel_dsl("one_compartment"))
    ll = eq.estimate_log_likelihood(ps.Subject("a", c["ops"]), c["params"], em)
    want = oracle.Model("one_compartment").log_likelihood(oracle.Subject(c["ops"]), c["params"], oracle.ErrorModels([c["error_model"]]))
    assert ll == pytest.approx(want, rel=1e-13)
    ode = ps.Equation.from_dsl("name = o\nkind = ode\nparams = ke, v\nstates = central\noutputs = outeq_0\nbolus(input_0) -> central\n"
                               "dx(central) = -ke * central\nout(outeq_0) = central / v ~ continuous()\n")
    assert math.exp(ode.estimate_log_likelihood(ps.Subject("a", c["ops"]), c["params"], em)) == pytest.approx(math.exp(want), rel=1e-2)


def test_particle_filter_fixture_is_finite(ps):
    # tests/test_pf.rs:8-59 asserts finiteness only
    c = FX.PF_TEST
    eq = ps.Equation.from_dsl(c["dsl"])
    em = ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(*c["error_model"][2]), 0.0))
    ll = eq.estimate_log_likelihood(ps.Subject("a", c["ops"]), c["params"], em)
    assert math.isfinite(ll)


# ---- error behaviour ------------------------------------------------------------------------------------------
def test_first_error_aborts_like_matrix_rs(ps, W, H):
    w = W.make("c1", nsub=5, nspp=40)
    eq, data, _ = H.product_objects(w)
    bad = ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(0.0, 0.0, 0.0, 0.0), 0.0))   # sigma == 0
    with pytest.raises(ps.PharmsolError) as e:
        ps.log_likelihood_matrix(eq, data, w["support_points"], bad)
    assert e.value.code in (1, 2, 3)
    # imaginary roots (replaces the reference's panic!): negative rate constant in a 2-cpt model
    eq2 = ps.Equation.from_dsl(FX.kernel_dsl("two_compartments"))
    spp = np.array([[0.1, 3.0, 1.0, 1.0], [1.0, -3.0, 1.5, 1.0], [0.2, 1.0, 1.0, 1.0]])
    em = ps.AssayErrorModels().add("outeq_0", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0))
    ops = [("bolus", 0.0, 100.0, "0"), ("observation", 1.0, 50.0, "0")]
    with pytest.raises(ps.PharmsolError) as e:
        ps.log_likelihood_matrix(eq2, ps.Data([ps.Subject("a", ops), ps.Subject("b", ops)]), spp, em)
    assert e.value.code == 12 and e.value.pair == 0 + 1 * 2     # first failing pair: subject 0, support point 1


def test_first_error_in_a_later_chunk_of_a_pipelined_host_call(ps):
    """Host calls are pipelined over column chunks (closed-form models: up to 16 equal chunks; adaptive ODE models:
    7/8 + 1/8 of the columns).  The first failing pair is reported with its GLOBAL index wherever it falls, the columns
    before and after it are still evaluated, and two failures report the smaller pair (matrix.rs:96-104)."""
    ops = [("bolus", 0.0, 100.0, "0")] + [("observation", float(t), 50.0 / t, "0") for t in (1, 2, 4)]
    nsub = 600
    data = ps.Data([ps.Subject(f"s{i}", ops) for i in range(nsub)])
    em = ps.AssayErrorModels().add("outeq_0", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0))
    # closed form: 600 x 2048 x 8 B = 9.8 MB -> 4 chunks of 512 columns
    eq = ps.Equation.from_dsl(FX.kernel_dsl("two_compartments"))
    spp = np.tile(np.array([[0.1, 3.0, 1.0, 1.0]]), (2048, 1)) * (1.0 + 1e-4 * np.arange(2048)[:, None])
    good = ps.log_likelihood_matrix(eq, data, spp, em)
    assert np.all(np.isfinite(good))
    for bad_cols in ([1500], [1900, 700], [2047]):
        bad = spp.copy()
        bad[bad_cols] = [1.0, -3.0, 1.5, 1.0]   # (ke + kcp + kpc)^2 < 4 ke kpc: imaginary roots
        with pytest.raises(ps.PharmsolError) as e:
            ps.log_likelihood_matrix(eq, data, bad, em)
        assert e.value.code == 12 and e.value.pair == 0 + min(bad_cols) * nsub
    # adaptive ODE: 600 x 8192 x 8 B = 39 MB -> two chunks (7168 + 1024 columns); a failure in the small tail chunk
    src = ("name = chunk_ode\nkind = ode\nparams = ke, v\nstates = central\noutputs = outeq_0\nbolus(input_0) -> central\n"
           "dx(central) = -ke * central\nout(outeq_0) = central / v ~ continuous()\n")
    eo = ps.Equation.from_dsl(src)
    spo = np.column_stack([0.1 + 1e-5 * np.arange(8192), np.full(8192, 2.0)])
    ok = ps.log_likelihood_matrix(eo, data, spo, em)
    assert np.all(np.isfinite(ok))
    spo[8000, 0] = float("nan")
    with pytest.raises(ps.PharmsolError) as e:
        ps.log_likelihood_matrix(eo, data, spo, em)
    assert e.value.pair == 0 + 8000 * nsub


def test_wrong_parameter_count_and_unknown_labels(ps, W, H):
    w = W.make("c1", nsub=2, nspp=4)
    eq, data, ems = H.product_objects(w)
    with pytest.raises(ps.PharmsolError) as e:
        ps.log_likelihood_matrix(eq, data, np.ones((4, 3)), ems)
    assert "expects 2 parameter" in str(e.value)
    bad = ps.Data([ps.Subject("x", [("bolus", 0.0, 1.0, "nonexistent"), ("observation", 1.0, 1.0, "cp")])])
    with pytest.raises(ps.PharmsolError) as e:
        ps.log_likelihood_matrix(eq, bad, w["support_points"], ems)
    assert e.value.code in (10, 13)
    bad = ps.Data([ps.Subject("x", [("infusion", 0.0, 1.0, "iv", 1.0), ("observation", 1.0, 1.0, "nope")])])
    with pytest.raises(ps.PharmsolError) as e:
        ps.log_likelihood_matrix(eq, bad, w["support_points"], ems)
    assert e.value.code == 11


def test_empty_and_ragged_inputs(ps, oracle, W, H):
    w = W.make("c1", nsub=3, nspp=5)
    eq, data, ems = H.product_objects(w)
    assert ps.log_likelihood_matrix(eq, data, np.empty((0, 2)), ems).shape == (3, 0)
    # ragged: subjects with 0, 1 and many observations; a subject without any dose
    subs = [("a", [("infusion", 0.0, 100.0, "iv", 1.0)]),
            ("b", [("observation", 1.0, 0.5, "cp")]),
            ("c", [("infusion", 0.0, 100.0, "iv", 1.0)] + [("observation", 0.25 * k, 1.0, "cp") for k in range(1, 60)]),
            ("d", [("infusion", 2.0, 100.0, "iv", 0.25), ("infusion", 2.1, 50.0, "iv", 3.0), ("observation", 2.2, 1.0, "cp"), ("observation", 9.0, 0.2, "cp")])]
    data = ps.Data([ps.Subject(i, o) for i, o in subs])
    psi = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    om = oracle.Model("one_cpt_iv")
    ref = om.log_likelihood_matrix(oracle.Data([oracle.Subject(o, i) for i, o in subs]), w["support_points"], oracle.ErrorModels([w["error_models"]["cp"]]))
    assert np.all(psi[0] == 0.0)          # no observations -> empty sum
    ll_close(psi, ref, 1, 1e-12)


def test_psi_is_exp_of_log_matrix(ps, W, H):
    import warnings
    w = W.make("c1", nsub=4, nspp=33)
    eq, data, ems = H.product_objects(w)
    lg = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p = ps.psi(eq, data, w["support_points"], ems)
    assert np.allclose(p, np.exp(lg), rtol=1e-14, atol=0)


# ---- module provenance: ahead-of-time vs NVRTC ---------------------------------------------------------------------
def test_user_model_goes_through_nvrtc_on_device(ps, tmp_path, monkeypatch):
    from pharmsol_b200 import _lib
    monkeypatch.setenv("PHARMSOL_B200_CUBIN_CACHE", str(tmp_path))
    src = "name = user_model_%d\nkind = ode\nparams = ke, v\nstates = central\noutputs = cp\nbolus(iv) -> central\n" \
          "dx(central) = -ke * central * 1.000001\nout(cp) = central / v ~ continuous()\n" % os.getpid()
    eq = ps.Equation.from_dsl(src).with_tolerances(1e-10, 1e-10)
    assert eq._model.compile(_lib.context(0)) == "nvrtc"
    got = gpu_predictions(ps, eq, [("bolus", 0.0, 100.0, "iv"), ("missing_observation", 2.0, "cp")], [0.3, 10.0])
    assert got[0] == pytest.approx(10.0 * math.exp(-0.3 * 1.000001 * 2.0), rel=1e-8)
    eq2 = ps.Equation.from_dsl(src)
    assert eq2._model.compile(_lib.context(0)) == "cubin-cache"
    from benches import workloads
    assert ps.Equation.from_dsl(workloads.model_source("c1_one_cpt_iv"))._model.compile(_lib.context(0)) == "aot"


# ---- BASELINE full sizes through size-independent properties -----------------------------------------------------------
def test_c1_full_size_properties(ps, oracle, H, W):
    """1,000 x 1,000: spot-check 300 random pairs against the oracle; permuting support points permutes
    columns bit-exactly; a subject subset reproduces its rows bit-exactly."""
    w = W.make("c1")
    eq, data, ems = H.product_objects(w)
    spp = w["support_points"]
    psi = ps.log_likelihood_matrix(eq, data, spp, ems)
    assert psi.shape == (1000, 1000) and np.all(np.isfinite(psi))
    om, od, oe = H.oracle_objects(w)
    rng = np.random.default_rng(3)
    for i, j in zip(rng.integers(0, 1000, 300), rng.integers(0, 1000, 300)):
        want = om.log_likelihood(od.subjects[i], spp[j], oe)
        assert abs(psi[i, j] - want) <= 1e-12 * (abs(want) + 10)
    perm = rng.permutation(1000)
    assert np.array_equal(ps.log_likelihood_matrix(eq, data, spp[perm], ems), psi[:, perm])
    rows = [3, 500, 999]
    sub = ps.Data([data.subjects[r] for r in rows])
    assert np.array_equal(ps.log_likelihood_matrix(eq, sub, spp, ems), psi[rows])


def test_c2_full_size_properties(ps, oracle, H, W):
    """500 x 20,000 Dopri5 at the bench tolerance: finite everywhere, spot-checked against the closed
    form, column-permutation invariant."""
    w = W.make("c2")
    eq, data, ems = H.product_objects(w)
    eq.with_solver(ps.OdeSolver.Dopri5).with_tolerances(1e-6, 1e-6)
    spp = w["support_points"]
    psi = ps.log_likelihood_matrix(eq, data, spp, ems)
    assert psi.shape == (500, 20000) and np.all(np.isfinite(psi))
    om, od, oe = H.oracle_objects(dict(w, oracle_model=w["oracle_truth_model"]))
    rng = np.random.default_rng(5)
    worst = 0.0
    for i, j in zip(rng.integers(0, 500, 200), rng.integers(0, 20000, 200)):
        want = om.log_likelihood(od.subjects[i], spp[j], oe)
        worst = max(worst, abs(psi[i, j] - want) / (abs(want) + 12))
    assert worst <= 1e-4, worst     # rtol = atol = 1e-6 solver: likelihood amplifies prediction error by |z| pred/sigma
    perm = rng.permutation(20000)[:4096]
    assert np.array_equal(ps.log_likelihood_matrix(eq, data, spp[perm], ems), psi[:, perm])


# ---- randomized timelines: every closed-form kernel vs the oracle -----------------------------------------------------------
def _random_subject(rng, absorb, n_occ):
    ops = []
    grid = np.round(rng.uniform(0.0, 48.0, 40) * 4) / 4          # quarter-hour grid -> plenty of exact ties
    for occ in range(n_occ):
        if occ:
            ops.append(("reset",))
        for _ in range(rng.integers(1, 5)):
            ops.append(("bolus", float(rng.choice(grid)), float(rng.uniform(10, 500)), "1" if (absorb and rng.random() < 0.4) else "0"))
        for _ in range(rng.integers(0, 4)):
            ops.append(("infusion", float(rng.choice(grid)), float(rng.uniform(10, 500)), "0", float(rng.choice([0.25, 0.5, 1.0, 3.0, 7.5]))))
        for _ in range(rng.integers(1, 12)):
            t = float(rng.choice(grid))
            ops.append(("missing_observation", t, "0") if rng.random() < 0.3 else ("observation", t, float(rng.uniform(0.1, 20.0)), "0"))
    return ops


@pytest.mark.parametrize("kernel", list(FX.KERNEL_PARAMS))
def test_random_timelines_match_oracle(ps, oracle, kernel):
    """Overlapping infusions, ties between observations / boluses / infusion boundaries, several occasions,
    missing observations, ragged subjects: predictions <= 1e-12 relative (floor 1e-9 of the largest amount),
    log-likelihood <= 1e-12 scaled."""
    seed = sum(kernel.encode()) + 20261018          # deterministic per kernel (hash() is salted per process)
    rng = np.random.default_rng(seed)
    absorb = kernel.endswith("with_absorption")
    subjects = [(f"r{i}", _random_subject(rng, absorb, int(rng.integers(1, 4)))) for i in range(12)]
    names = FX.KERNEL_PARAMS[kernel]
    nspp = 64

    def draw(name):
        if name in ("v", "vc", "vp", "v2", "v3"):
            return rng.uniform(5.0, 80.0, nspp)
        if name in ("cl", "q", "q2", "q3"):
            return rng.uniform(0.5, 20.0, nspp)
        return rng.uniform(0.02, 2.5, nspp)
    spp = np.stack([draw(n) for n in names], axis=1)
    eq = ps.Equation.from_dsl(FX.kernel_dsl(kernel))
    data = ps.Data([ps.Subject(i, o) for i, o in subjects])
    em = ("additive", 0.05, (0.1, 0.15, 0.0, 0.0))
    ems = ps.AssayErrorModels().add("outeq_0", ps.AssayErrorModel.additive(ps.ErrorPoly(*em[2]), em[1]))
    psi = ps.log_likelihood_matrix(eq, data, spp, ems)
    om = oracle.Model(kernel)
    od = oracle.Data([oracle.Subject(o, i) for i, o in subjects])
    ref = om.log_likelihood_matrix(od, spp, oracle.ErrorModels([em]))
    nobs = max(sum(1 for o in ops if o[0] == "observation") for _, ops in subjects)
    ll_close(psi, ref, nobs, 1e-11 if kernel.startswith("three") else 1e-12)
    pred, offs = eq.predictions_matrix(data, spp[:8])
    for i in range(len(subjects)):
        for j in range(8):
            want = om.predictions(od.subjects[i], spp[j])
            got = pred[offs[i]:offs[i + 1], j]
            assert got.shape == want.shape
            floor = 1e-9 * max(1.0, np.max(np.abs(want)))
            assert rel(got, want, floor).max() <= (2e-11 if kernel.startswith("three") else 1e-12), (kernel, i, j)


def test_work_balanced_column_order_does_not_change_psi(ps, H, W, monkeypatch):
    """ODE launches probe per-column step counts on a few subjects and evaluate the columns in sorted order
    (work-balanced warps).  psi must be bit-identical with the balancing switched off."""
    w = W.make("c2", nsub=64, nspp=2048)
    eq, data, ems = H.product_objects(w)
    eq.with_solver(ps.OdeSolver.Dopri5).with_tolerances(1e-6, 1e-6)
    monkeypatch.setenv("PHARMSOL_B200_BALANCE", "0")
    plain = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    monkeypatch.setenv("PHARMSOL_B200_BALANCE", "1")
    balanced = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    assert np.array_equal(plain, balanced)
    from pharmsol_b200 import _lib
    assert _lib.context(0).last_counters["steps"] > 0


@pytest.mark.parametrize("solver,tol,bar", [("Dopri5", 1e-10, 1e-6), ("Tsit45", 1e-10, 1e-6), ("Rodas4", 1e-9, 2e-6), ("Sdirk4", 1e-9, 2e-5)])
def test_dsl_feature_model_vs_scipy_golden(ps, solver, tol, bar):
    """DSL front end -> CUDA C on a model that uses array states, constants, statement-level if / else-if,
    a conditional expression, intrinsics, explicit rate(), a time-dependent right-hand side (covariate and t),
    lag and fa, against an independent SciPy DOP853 integration (scripts/gen_golden.py)."""
    g = golden("dsl_features")
    eq = ps.Equation.from_dsl(g["dsl"]).with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(tol, tol)
    for c in g["cases"]:
        got = gpu_predictions(ps, eq, [tuple(o) for o in c["ops"]], c["params"])
        err = rel(got, c["predictions"], 1e-4).max()
        assert err <= bar, (solver, c["params"], err)     # SDIRK4: the simplified-Newton stopping rule limits it to ~1e-5 here


# ---- log_likelihood_batch + ResidualErrorModels -----------------------------------------------------
@pytest.mark.parametrize("name,tol", [("c1", 1e-12), ("c2", 1e-6)])
def test_log_likelihood_batch_vs_oracle(ps, oracle, H, W, name, tol):
    """likelihood/mod.rs:119-177: subject i with parameter row i, prediction-based sigma (data/residual_error.rs)."""
    w = W.make(name, nsub=97, nspp=97)
    eq, data, _ = H.product_objects(w)
    okw = {}
    if name == "c2":
        eq.with_solver(ps.OdeSolver.Dopri5).with_tolerances(1e-10, 1e-10)
        okw = dict(solver="dopri5", rtol=1e-12, atol=1e-12)
    om, od, _ = H.oracle_objects(w, **okw)
    prm = w["support_points"]
    for model, omodel in [(ps.ResidualErrorModel.constant(0.5), ("constant", 0.5, 0.0)),
                          (ps.ResidualErrorModel.proportional(0.15), ("proportional", 0.0, 0.15)),
                          (ps.ResidualErrorModel.combined(0.1, 0.2), ("combined", 0.1, 0.2)),
                          (ps.ResidualErrorModel.exponential(0.3), ("exponential", 0.3, 0.0))]:
        got = ps.log_likelihood_batch(eq, data, prm, ps.ResidualErrorModels().add(0, model))
        want = oracle.log_likelihood_batch(om, od, prm, [omodel])
        assert got.shape == (97,) and np.all(np.isfinite(got))
        assert np.max(np.abs(got - want) / (np.abs(want) + 12)) <= tol
    # no model for the output -> -inf for everyone; wrong row count -> error (mod.rs:128-134)
    assert np.all(np.isneginf(ps.log_likelihood_batch(eq, data, prm, ps.ResidualErrorModels())))
    with pytest.raises(ps.PharmsolError) as e:
        ps.log_likelihood_batch(eq, data, prm[:5], ps.ResidualErrorModels().add(0, ps.ResidualErrorModel.constant(1.0)))
    assert "rows but there are 97 subjects" in str(e.value)


def test_log_likelihood_batch_failed_simulation_is_neg_inf(ps):
    eq = ps.Equation.from_dsl(FX.kernel_dsl("two_compartments"))
    ops = [("bolus", 0.0, 100.0, "0"), ("observation", 1.0, 50.0, "0")]
    data = ps.Data([ps.Subject("a", ops), ps.Subject("b", ops)])
    prm = np.array([[0.1, 3.0, 1.0, 1.0], [1.0, -3.0, 1.5, 1.0]])      # second row: imaginary roots
    out = ps.log_likelihood_batch(eq, data, prm, ps.ResidualErrorModels().add(0, ps.ResidualErrorModel.constant(1.0)))
    assert np.isfinite(out[0]) and np.isneginf(out[1])


def test_c3_full_shard_properties(ps, oracle, H, W):
    """C3 at the per-GPU shard of BASELINE configs[2] (10,000 subjects x 6,250 support points): finite, spot-checked
    against the oracle, bit-exact under column permutation and subject subsetting."""
    w = W.make("c3", nsub=10000, nspp=6250)
    eq, data, ems = H.product_objects(w)
    spp = w["support_points"]
    psi = ps.log_likelihood_matrix(eq, data, spp, ems)
    assert psi.shape == (10000, 6250) and np.all(np.isfinite(psi))
    om, od, oe = H.oracle_objects(w)
    rng = np.random.default_rng(11)
    worst = 0.0
    for i, j in zip(rng.integers(0, 10000, 200), rng.integers(0, 6250, 200)):
        want = om.log_likelihood(od.subjects[i], spp[j], oe)
        worst = max(worst, abs(psi[i, j] - want) / (abs(want) + 10))
    assert worst <= 1e-10, worst
    perm = rng.permutation(6250)[:512]
    sub = ps.Data([data.subjects[r] for r in (0, 4999, 9999)])
    assert np.array_equal(ps.log_likelihood_matrix(eq, sub, spp[perm], ems), psi[[0, 4999, 9999]][:, perm])


def test_c4_full_size_properties(ps, oracle, H, W):
    """C4 at BASELINE size (2,000 x 10,000, RODAS4 at the bench tolerance): finite, spot-checked against a tight
    explicit-RK oracle run, bit-exact under column permutation (which also changes the work-balanced order)."""
    w = W.make("c4")
    eq, data, ems = H.product_objects(w)
    eq.with_solver(ps.OdeSolver.Rodas4).with_tolerances(1e-6, 1e-6)
    spp = w["support_points"]
    psi = ps.log_likelihood_matrix(eq, data, spp, ems)
    assert psi.shape == (2000, 10000) and np.all(np.isfinite(psi))
    om, od, oe = H.oracle_objects(w, solver="dopri5", rtol=1e-10, atol=1e-10)
    rng = np.random.default_rng(13)
    worst = 0.0
    for i, j in zip(rng.integers(0, 2000, 150), rng.integers(0, 10000, 150)):
        want = om.log_likelihood(od.subjects[i], spp[j], oe)
        worst = max(worst, abs(psi[i, j] - want) / (abs(want) + 8))
    assert worst <= 2e-4, worst     # rtol = atol = 1e-6 solver; the likelihood amplifies prediction error by |z| pred/sigma
    perm = rng.permutation(10000)[:2048]
    assert np.array_equal(ps.log_likelihood_matrix(eq, data, spp[perm], ems), psi[:, perm])


def test_c5_reference_stepper_properties(ps, H, W):
    """C5 (reference adaptive EM + particle filter, 1,000 particles) on a slice of the BASELINE population: the result
    of a pair depends only on (seed, pair index within the call, parameters), so re-running is bit-identical and
    -inf (not NaN) marks pairs whose particle likelihood underflows (sde/mod.rs:699-703)."""
    w = W.make("c5", nsub=8, nspp=64, particles=1000)
    eq, data, ems = H.product_objects(w)
    eq.with_particles(1000).with_mode(ps.SdeMode.ParticleFilter).with_seed(42)
    a = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    b = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    assert np.array_equal(a, b, equal_nan=True) and not np.any(np.isnan(a)) and not np.any(np.isposinf(a))
    assert np.isfinite(a).mean() > 0.2


def test_host_call_pipelines_chunks_without_changing_psi(ps, H, W):
    """The host-buffer call splits large outputs into column chunks (compute chunk k+1 while chunk k is copied back).
    Result must equal the single-launch device-resident path bit for bit, including the first-error pair index."""
    import torch
    w = W.make("c1", nsub=3000, nspp=6000)           # 144 MB of psi -> 4 chunks
    eq, data, ems = H.product_objects(w)
    host = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    job = ps.ResidentPsi(eq, data, w["support_points"], ems, shard=False)
    job.launch()
    resident = job.finish().cpu().numpy()
    assert np.array_equal(host, resident)
    # an error in a late chunk still reports its GLOBAL pair index
    eq2 = ps.Equation.from_dsl(FX.kernel_dsl("two_compartments"))
    ops = [("bolus", 0.0, 100.0, "0")] + [("observation", float(t), 50.0, "0") for t in range(1, 4)]
    data2 = ps.Data([ps.Subject(f"s{i}", ops) for i in range(2000)])
    spp = np.tile(np.array([[0.1, 3.0, 1.0, 1.0]]), (9000, 1))
    spp[8123] = [1.0, -3.0, 1.5, 1.0]                  # imaginary roots
    em = ps.AssayErrorModels().add("outeq_0", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0))
    with pytest.raises(ps.PharmsolError) as e:
        ps.log_likelihood_matrix(eq2, data2, spp, em)
    assert e.value.code == 12 and e.value.pair == 0 + 8123 * 2000


# ---- reference tests with LITERAL expected values for the DSL path ----------------------------------------------------
def test_dsl_numeric_equality_literals(ps):
    """tests/dsl_numeric_equality.rs:13-81: real-valued covariate equality selects branches; exact expected outputs."""
    src = """
name = numeric_equality
kind = ode

params = a, b, c
covariates = drug
states = central
outputs = e, non_integral, not_one

dx(central) = 0

out(e) = if (drug == 1) a else if (drug == 2) b else c
out(non_integral) = if (drug == 1.1) a else if (drug == 2.1) b else c
out(not_one) = if (drug != 1) a else b
"""
    eq = ps.Equation.from_dsl(src)

    def value(drug, output):
        ops = [("covariate", "drug", 0.0, drug), ("missing_observation", 0.0, output)]
        return gpu_predictions(ps, eq, ops, [10.0, 20.0, 30.0])[0]
    for drug, want in [(1.0, 10.0), (2.0, 20.0), (3.0, 30.0)]:
        assert value(drug, "e") == want
    for drug, want in [(1.0, 30.0), (2.0, 30.0), (3.0, 30.0), (1.1, 10.0), (2.1, 20.0), (3.1, 30.0)]:
        assert value(drug, "non_integral") == want
    for drug, want in [(1.0, 20.0), (2.0, 10.0), (3.0, 10.0), (1.1, 10.0), (2.1, 10.0), (3.1, 10.0)]:
        assert value(drug, "not_one") == want


@pytest.mark.parametrize("keyword", ["t", "time"])
def test_dsl_time_keyword(ps, keyword):
    """tests/dsl_time_keyword.rs:10-80: `t` / `time` is the current simulation time in outputs."""
    src = f"""
name = time_probe
kind = ode

params = ke
states = central
outputs = cp, time_echo

infusion(iv) -> central

dx(central) = -ke * central

out(cp) = central
out(time_echo) = {keyword}
"""
    eq = ps.Equation.from_dsl(src)
    times = [0.5, 1.0, 2.5, 4.0]
    ops = [("infusion", 0.0, 100.0, "iv", 1.0)] + [("missing_observation", t, "time_echo") for t in times]
    got = gpu_predictions(ps, eq, ops, [1.0])
    assert len(got) == 4 and np.max(np.abs(got - np.array(times))) < 1e-6


def test_concurrent_host_threads_share_one_equation(ps, H, W):
    """`Equation: Sync` (equation/mod.rs:377): one model / population / context used from several host threads at once
    (ctypes releases the GIL, the library serialises per context) gives the serial results."""
    import threading
    w = W.make("c2", nsub=24, nspp=512)
    eq, data, ems = H.product_objects(w)
    eq.with_solver(ps.OdeSolver.Dopri5).with_tolerances(1e-6, 1e-6)
    spp = w["support_points"]
    serial = [ps.log_likelihood_matrix(eq, data, spp[k * 128:(k + 1) * 128], ems) for k in range(4)]
    out, errs = [None] * 4, []

    def work(k):
        try:
            for _ in range(3):
                out[k] = ps.log_likelihood_matrix(eq, data, spp[k * 128:(k + 1) * 128], ems)
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for k in range(4):
        assert np.array_equal(out[k], serial[k])
    # the concurrent callers did not queue behind one lock: the context grew extra stream lanes
    assert eq._ctx().num_lanes >= 2


def test_concurrent_callers_with_errors_keep_their_own_status(ps):
    """Lanes carry their own error word: a failing matrix on one thread must not leak its status into a clean matrix
    evaluated at the same time on another thread (matrix.rs:96-104 is per call)."""
    import threading
    eq = ps.Equation.from_dsl(FX.kernel_dsl("two_compartments"))
    ops = [("bolus", 0.0, 100.0, "0"), ("observation", 1.0, 50.0, "0"), ("observation", 2.0, 30.0, "0")]
    data = ps.Data([ps.Subject(f"s{i}", ops) for i in range(64)])
    ems = ps.AssayErrorModels().add("outeq_0", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0))
    good = np.tile(np.array([[0.1, 3.0, 1.0, 1.0]]), (700, 1))
    bad = good.copy()
    bad[555] = [1.0, -3.0, 1.5, 1.0]        # imaginary roots
    ref = ps.log_likelihood_matrix(eq, data, good, ems)
    res = {}

    def run(name, spp):
        for _ in range(20):
            try:
                res[name] = ("ok", ps.log_likelihood_matrix(eq, data, spp, ems))
            except ps.PharmsolError as e:
                res[name] = ("err", e.code, e.pair)
    ts = [threading.Thread(target=run, args=("good", good)), threading.Thread(target=run, args=("bad", bad))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert res["good"][0] == "ok" and np.array_equal(res["good"][1], ref)
    assert res["bad"] == ("err", 12, 555 * 64)


def test_solver_failure_is_a_status_not_a_hang(ps):
    """A finite-time blow-up (dx = x^2) cannot be integrated: every solver must give PharmsolError(SolverFailure = 7)
    for the failing pair (DiffsolError in the reference) and keep the healthy pair finite."""
    src = "name = blowup\nkind = ode\nparams = k, v\nstates = x\noutputs = cp\nbolus(iv) -> x\ndx(x) = k * x * x\nout(cp) = x / v ~ continuous()\n"
    ops = [("bolus", 0.0, 1.0, "iv"), ("observation", 0.5, 1.0, "cp"), ("observation", 3.0, 1.0, "cp")]
    em = ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0))
    spp = np.array([[0.1, 1.0], [1.0, 1.0]])        # k = 1 blows up at t = 1
    for solver in ("Dopri5", "Tsit45", "Rodas4", "Sdirk4", "TrBdf2"):
        eq = ps.Equation.from_dsl(src).with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(1e-6, 1e-6).with_max_steps(20000)
        with pytest.raises(ps.PharmsolError) as e:
            ps.log_likelihood_matrix(eq, ps.Data([ps.Subject("a", ops)]), spp, em)
        assert e.value.code == 7 and e.value.pair == 1, (solver, e.value.code, e.value.pair)
        ok = ps.log_likelihood_matrix(eq, ps.Data([ps.Subject("a", ops)]), spp[:1], em)
        assert np.isfinite(ok).all()


def test_more_subjects_than_grid_rows(ps, oracle):
    """nsub > 65535 (the grid's y extent): the CTA loops over subjects."""
    n = 66000
    rng = np.random.default_rng(5)
    amt = rng.uniform(100, 600, n)
    subs = [(f"s{i}", [("infusion", 0.0, float(amt[i]), "iv", 0.5), ("observation", 1.0, 2.0, "cp"), ("observation", 6.0, 0.7, "cp")]) for i in range(n)]
    from benches import workloads
    eq = ps.Equation.from_dsl(workloads.model_source("c1_one_cpt_iv"))
    data = ps.Data([ps.Subject(i, o) for i, o in subs])
    em = ("additive", 0.0, (0.1, 0.1, 0.0, 0.0))
    spp = np.array([[0.3, 100.0], [0.7, 60.0], [0.1, 200.0]])
    psi = ps.log_likelihood_matrix(eq, data, spp, ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(*em[2]), 0.0)))
    assert psi.shape == (n, 3) and np.all(np.isfinite(psi))
    om = oracle.Model("one_cpt_iv")
    oe = oracle.ErrorModels([em])
    for i in (0, 1, 65534, 65535, 65536, n - 1):
        for j in range(3):
            want = om.log_likelihood(oracle.Subject(subs[i][1]), spp[j], oe)
            assert abs(psi[i, j] - want) <= 1e-12 * (abs(want) + 2)


def test_particle_filter_converges_to_the_kalman_likelihood(ps, W):
    """Independent truth for the SDE path: the C5 model is linear-Gaussian (dx = -ke x + rate, additive diffusion sigma,
    Gaussian assay error), so the exact likelihood is a Kalman filter.  With small fixed EM steps and many particles the
    particle-filter estimate (log of the seed-averaged likelihood) must converge to it.  Euler-Maruyama and the reference's
    closed-interval drift-rate rule `start <= t <= start + dur` (one extra step of infusion) are O(dt) biased, so the test
    checks first-order convergence: the bias at dt = 0.001 is small and ~10x smaller than at dt = 0.01."""
    amt, dur, v_true = 500.0, 1.0, 100.0
    t_obs = [0.5, 1.0, 2.0, 3.0, 4.0, 6.0, 8.0, 12.0]
    rng = np.random.default_rng(21)
    y = [float(amt / dur / 0.4 * (1 - math.exp(-0.4 * min(t, dur))) * math.exp(-0.4 * max(t - dur, 0.0)) / v_true * math.exp(0.1 * rng.standard_normal())) for t in t_obs]
    ops = [("infusion", 0.0, amt, "iv", dur)] + [("observation", t, yy, "cp") for t, yy in zip(t_obs, y)]
    spp = np.array([[0.4, 2.0, 100.0], [0.3, 5.0, 90.0], [0.6, 1.0, 120.0], [0.4, 8.0, 100.0]])      # ke, sigma, v
    c0, c1 = 0.1, 0.1

    def kalman(ke, sig, v):
        m, P, t, ll = 0.0, 0.0, 0.0, 0.0
        for tk, yk in zip(t_obs, y):
            # propagate across [t, tk], splitting at the infusion end
            for (a, b) in ([(t, min(tk, dur)), (min(tk, dur), tk)] if t < dur < tk else [(t, tk)]):
                if b <= a:
                    continue
                dt, r = b - a, (amt / dur if b <= dur else 0.0)
                e = math.exp(-ke * dt)
                m = m * e + r / ke * (1 - e)
                P = P * e * e + sig * sig * (1 - e * e) / (2 * ke)
            t = tk
            s = c0 + c1 * yk                                  # additive error model: sigma from the observation
            S = P / (v * v) + s * s
            innov = yk - m / v
            ll += -0.5 * math.log(2 * math.pi * S) - 0.5 * innov * innov / S
            K = (P / v) / S
            m, P = m + K * innov, P - K * P / v
        return ll
    data = ps.Data([ps.Subject("a", ops)])
    ems = ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(c0, c1, 0.0, 0.0), 0.0))
    exact = np.array([kalman(*p) for p in spp])
    nseed = 16

    def estimate(dt):
        eq = ps.Equation.from_dsl(W.model_source("c5_one_cpt_sde"))
        eq.with_particles(8192).with_mode(ps.SdeMode.ParticleFilter).with_stepper(ps.EmMode.FixedStep, dt)
        ll = np.stack([ps.log_likelihood_matrix(eq.with_seed(7000 + s), data, spp, ems)[0] for s in range(nseed)])     # (nseed, 4)
        shift = ll.max(axis=0)
        lik = np.exp(ll - shift)
        return np.log(lik.mean(axis=0)) + shift, lik.std(axis=0, ddof=1) / math.sqrt(nseed) / lik.mean(axis=0)
    est_c, se_c = estimate(0.01)
    est_f, se_f = estimate(0.001)
    err_c, err_f = est_c - exact, est_f - exact
    # converged: within 0.15 log-units (of |ll| up to 20) at dt = 0.001, Monte-Carlo error negligible
    assert np.all(np.abs(err_f) <= 0.15 + 4 * se_f), (est_f, exact, se_f)
    # first order in dt: a 10x smaller step gives a ~10x smaller bias for every support point
    ratio = err_c / err_f
    assert np.all((ratio > 6.0) & (ratio < 15.0)), (err_c, err_f, ratio)


@pytest.mark.parametrize("solver,tol,bar", [("Rodas4", 1e-9, 1e-5), ("Dopri5", 1e-10, 1e-5), ("Sdirk4", 1e-9, 1e-4)])
def test_stiff_hybrid_phage_model_vs_radau_golden(ps, solver, tol, bar):
    """The reference's long-horizon stiff regression model (ode/mod.rs:1460-1600: 6 states, soft-plus clamps, logistic
    growth to 1e10, 25 infusions of 1e9-3e9 units over 1.25e-3 h) through DSL -> CUDA, against SciPy Radau (rtol 1e-11)
    on 17 observation times x 2 outputs.  Exercises the restart logic at 50 infusion boundaries and the register LU at N = 6."""
    g = golden("hybrid_phage")
    eq = ps.Equation.from_dsl(g["dsl"]).with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(tol, tol)
    got = gpu_predictions(ps, eq, [tuple(o) for o in g["ops"]], g["params"])
    want = np.array(g["predictions"])
    assert got.shape == want.shape and np.all(np.isfinite(got))
    err = rel(got, want, 1.0)      # amounts are 1e4 .. 3e9; floor of 1 unit
    assert err.max() <= bar, (solver, err.max(), int(err.argmax()))


@pytest.mark.parametrize("solver", ["Dopri5", "Tsit45", "Rodas4", "Sdirk4", "TrBdf2", "Bdf", "Esdirk34"])
def test_ode_event_times_within_ulps_of_boundaries(ps, solver):
    """ode/mod.rs:1347-1455: (1) an infusion that ends one ULP after an observation — the loop must keep integrating to the
    next observation (expected value in closed form); (2) observations 16 ULPs on either side of a bolus time must not
    error and give 4 predictions."""
    src = "name = ulp1\nkind = ode\nparams = k\nstates = central\noutputs = cp\ninfusion(iv) -> central\ndx(central) = -k * central\nout(cp) = central ~ continuous()\n"
    eq = ps.Equation.from_dsl(src).with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(1e-8, 1e-8)
    dur = float(np.nextafter(10.0, 11.0)) - 5.0
    got = gpu_predictions(ps, eq, [("infusion", 5.0, 100.0, "iv", dur), ("observation", 10.0, 0.0, "cp"), ("observation", 20.0, 0.0, "cp")], [0.5])
    delivered = 100.0 * (1.0 - math.exp(-2.5)) / 2.5
    assert got[1] == pytest.approx(delivered * math.exp(-5.0), rel=1e-3)
    assert got[0] == pytest.approx(delivered, rel=1e-3)
    src2 = "name = ulp2\nkind = ode\nparams = k\nstates = central\noutputs = cp\nbolus(input_0) -> central\ndx(central) = -k * central\nout(cp) = central ~ continuous()\n"
    eq2 = ps.Equation.from_dsl(src2).with_solver(getattr(ps.OdeSolver, solver)).with_tolerances(1e-8, 1e-8)
    ulp = float(np.nextafter(12.0, 13.0)) - 12.0
    ops = [("bolus", 0.0, 200.0, "0"), ("bolus", 12.0, 100.0, "0"), ("missing_observation", 0.0, "cp"), ("missing_observation", 12.0 - 16.0 * ulp, "cp"),
           ("missing_observation", 12.0 + 16.0 * ulp, "cp"), ("missing_observation", 24.0, "cp")]
    got = gpu_predictions(ps, eq2, ops, [0.3])
    assert len(got) == 4 and np.all(np.isfinite(got))
    before = 200.0 * math.exp(-0.3 * 12.0)
    bar = 2e-5 if solver in ("TrBdf2", "Bdf", "Esdirk34") else 1e-6          # TR-BDF2 is second order: ~5e-6 at tol 1e-8
    assert got[0] == 0.0 and got[1] == pytest.approx(before, rel=bar) and got[2] == pytest.approx(before + 100.0, rel=bar)
    assert got[3] == pytest.approx((before + 100.0) * math.exp(-0.3 * 12.0), rel=bar)


@pytest.mark.parametrize("kernel", ["one_compartment", "one_compartment_with_absorption", "two_compartments_with_absorption"])
def test_ode_twin_matches_analytical_on_random_dosing(ps, kernel):
    """tests/ode_optimizations.rs:205-1100 (15 dosing scenarios: single / multiple boluses, oral absorption, overlapping
    infusions, bolus at an observation time, very fast / slow elimination, rapid absorption) generalised: the ODE twin of a
    closed-form kernel on randomized timelines, both on the device; reference tolerance 1e-2 relative, here 1e-6."""
    seed = sum(kernel.encode())
    rng = np.random.default_rng(seed)
    absorb = kernel.endswith("with_absorption")
    subjects = [(f"r{i}", _random_subject(rng, absorb, int(rng.integers(1, 3)))) for i in range(6)]
    twin = {
        "one_compartment": ("params = ke, v\nstates = central\n", "bolus(input_0) -> central\ninfusion(input_0) -> central\n",
                            "dx(central) = -ke * central\n"),
        "one_compartment_with_absorption": ("params = ka, ke, v\nstates = gut, central\n", "bolus(input_0) -> gut\nbolus(input_1) -> central\ninfusion(input_0) -> central\n",
                                            "dx(gut) = -ka * gut\ndx(central) = ka * gut - ke * central\n"),
        "two_compartments_with_absorption": ("params = ke, ka, kcp, kpc, v\nstates = gut, central, peripheral\n",
                                             "bolus(input_0) -> gut\nbolus(input_1) -> central\ninfusion(input_0) -> central\n",
                                             "dx(gut) = -ka * gut\ndx(central) = ka * gut - (ke + kcp) * central + kpc * peripheral\n"
                                             "dx(peripheral) = kcp * central - kpc * peripheral\n"),
    }[kernel]
    ode = ps.Equation.from_dsl(f"name = twin_{kernel}\nkind = ode\n{twin[0]}outputs = outeq_0\n{twin[1]}{twin[2]}out(outeq_0) = central / v ~ continuous()\n")
    ode.with_solver(ps.OdeSolver.Dopri5).with_tolerances(1e-10, 1e-10)
    ana = ps.Equation.from_dsl(FX.kernel_dsl(kernel))
    names = FX.KERNEL_PARAMS[kernel]
    base = {"ke": [0.1, 5.0, 0.001, 0.3], "ka": [1.0, 10.0, 0.5, 2.0], "kcp": [0.3, 1.0, 0.05, 0.2], "kpc": [0.2, 0.5, 0.02, 0.1], "v": [50.0, 10.0, 100.0, 25.0]}
    spp = np.array([[base[n][k] for n in names] for k in range(4)])      # nominal, very fast, very slow, mixed
    data = ps.Data([ps.Subject(i, o) for i, o in subjects])
    pa, offs = ana.predictions_matrix(data, spp)
    po, _ = ode.predictions_matrix(data, spp)
    scale = np.maximum(np.abs(pa), 1e-6 * np.abs(pa).max())
    assert (np.abs(po - pa) / scale).max() <= 1e-6


def test_macro_full_feature_analytical_parity(ps, oracle):
    """tests/full_feature_macro_parity.rs:200-352: the `analytical!` full-feature model (derived ke from two covariates
    feeding the kernel, lag, fa, init, covariate-scaled output).  The macro evaluates `derive` for the kernel at t = dt
    (SURVEY F5), which the device reproduces with CovTime.IntervalLength; the twin is the handwritten closure set."""
    src = """
name = macro_full
kind = analytical
params = ka, ke0, v, tlag, f_oral, base_gut, base_central
covariates = wt@linear, renal@linear
derived = ke, adjusted_v
states = gut, central
outputs = cp
bolus(oral) -> gut
bolus(load) -> central
infusion(iv) -> central
lag(oral) = tlag * sqrt(wt / 70.0) * pow(90.0 / renal, 0.1)
fa(oral) = min(max(f_oral * pow(renal / 90.0, 0.1), 0.0), 1.0)
ke = ke0 * pow(wt / 70.0, 0.75) * pow(renal / 90.0, 0.25)
adjusted_v = v * (wt / 70.0) * (1.0 + 0.001 * (renal - 90.0))
structure = one_compartment_with_absorption
init(gut) = base_gut + 0.03 * wt
init(central) = base_central + 0.08 * renal
out(cp) = central / adjusted_v ~ continuous()
"""
    _, _, p, ops, _ = FX.CORPUS["analytical_full"]
    eq = ps.Equation.from_dsl(src).with_cov_time(ps.CovTime.IntervalLength)
    got = gpu_predictions(ps, eq, ops, p)
    want = oracle.Model("macro_analytical_full").predictions(oracle.Subject(ops), p)
    assert rel(got, want, 1e-10).max() <= 1e-12          # reference's own bar: 1e-10
    # and the DSL-runtime semantics (derive at the sub-interval end) differ measurably on this fixture
    other = gpu_predictions(ps, ps.Equation.from_dsl(src), ops, p)
    assert rel(other, want, 1e-10).max() > 1e-6


def test_assay_error_model_sigma_anchors_on_device(ps, oracle):
    """data/error_model.rs:1186-1239 literal sigmas (sqrt(26), 2) and the per-observation ErrorPoly override, through the
    device likelihood."""
    eq = ps.Equation.from_dsl(FX.kernel_dsl("one_compartment"))
    p = [0.2, 1.0]
    s = ps.Subject("a", [("bolus", 0.0, 12.0, "0"), ("observation", 1.0, 20.0, "0")])
    pred = 12.0 * math.exp(-0.2)
    for model, sigma in ((ps.AssayErrorModel.additive(ps.ErrorPoly(1.0, 0.0, 0.0, 0.0), 5.0), math.sqrt(26.0)),
                         (ps.AssayErrorModel.proportional(ps.ErrorPoly(1.0, 0.0, 0.0, 0.0), 2.0), 2.0)):
        ll = eq.estimate_log_likelihood(s, p, ps.AssayErrorModels().add("outeq_0", model))
        assert ll == pytest.approx(oracle.lognormpdf(20.0, pred, sigma), rel=1e-13)
    s2 = ps.Subject("b", [("bolus", 0.0, 12.0, "0"), ("observation_with_error", 1.0, 20.0, "0", (0.5, 0.1, 0.0, 0.0), "none")])
    ll = eq.estimate_log_likelihood(s2, p, ps.AssayErrorModels().add("outeq_0", ps.AssayErrorModel.additive(ps.ErrorPoly(1.0, 0.0, 0.0, 0.0), 0.0)))
    assert ll == pytest.approx(oracle.lognormpdf(20.0, pred, 2.5), rel=1e-13)


def test_constant_and_fixed_covariates_in_the_kernel_parameters(ps, oracle, H, W):
    """A covariate with one record, or a fixed (carry-forward) covariate, feeding the kernel parameters of the
    3-compartment model: same likelihoods as the oracle."""
    w = W.make("c3", nsub=6, nspp=64)
    subs = []
    for k, (sid, ops) in enumerate(w["subjects"]):
        cov = [o for o in ops if o[0] == "covariate"]
        rest = [o for o in ops if o[0] != "covariate"]
        keep = cov[:1] if k % 2 == 0 else cov + [("covariate_fixed", 0, "wt", True)]     # single record | fixed (carry-forward)
        subs.append((sid, keep + rest))
    w = dict(w, subjects=subs)
    eq, data, ems = H.product_objects(w)
    psi = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    om, od, oe = H.oracle_objects(w)
    ref = om.log_likelihood_matrix(od, w["support_points"], oe)
    ll_close(psi, ref, 10, 1e-10)


def test_estimate_predictions_returns_prediction_records(ps):
    """Equation::estimate_predictions -> SubjectPredictions of Prediction{time, observation, prediction, outeq, occasion,
    censoring} (likelihood/prediction.rs:18-27, equation/mod.rs:526-532): rows follow the event order of each occasion."""
    from benches import workloads
    eq = ps.Equation.from_dsl(workloads.model_source("c4_mm_effect"))
    s = (ps.Subject.builder("p").bolus(0.0, 200.0, "load").observation(2.0, 3.5, "effect").missing_observation(1.0, "cp")
         .censored_observation(4.0, 0.2, "cp", ps.Censor.BLOQ).reset().infusion(0.0, 100.0, "iv", 1.0).observation(0.5, 1.0, "cp").build())
    sp = eq.estimate_predictions(s, [30.0, 2.0, 30.0, 5.0, 100.0, 3.0])
    rows = sp.predictions()
    assert len(sp) == 4
    assert [r.time() for r in rows] == [1.0, 2.0, 4.0, 0.5]
    assert [r.outeq() for r in rows] == [0, 1, 0, 0] and [r.occasion() for r in rows] == [0, 0, 0, 1]
    assert rows[0].observation() is None and rows[1].observation() == 3.5 and rows[2].censoring() == ps.Censor.BLOQ
    assert all(math.isfinite(r.prediction()) and r.prediction() > 0 for r in rows)
    assert sp.flat_predictions() == [r.prediction() for r in rows]


def test_sde_with_lag_fa_and_zero_noise_tracks_the_ode(ps, oracle):
    """SDE event handling (lag, bioavailability, bolus destinations, infusion rates in the drift) with the diffusion switched
    off: Euler-Maruyama at its 1e-2 tolerance must track the closed form of the same model."""
    sde = """
name = sde_lag
kind = sde
params = ka, ke, v, tlag, f_oral, s
states = depot, central
outputs = cp
particles = 8
bolus(oral) -> depot
infusion(iv) -> central
lag(oral) = tlag
fa(oral) = f_oral
dx(depot) = -ka * depot
dx(central) = ka * depot - ke * central
noise(central) = s
out(cp) = central / v ~ continuous()
"""
    ops = [("bolus", 0.0, 100.0, "oral"), ("infusion", 2.0, 50.0, "iv", 1.5), ("bolus", 6.0, 80.0, "oral")] + \
          [("missing_observation", float(t), "cp") for t in (0.25, 0.5, 1.0, 2.0, 3.0, 3.5, 5.0, 6.5, 7.0, 9.0, 12.0)]
    p = [1.2, 0.25, 20.0, 0.5, 0.8, 0.0]
    got = gpu_predictions(ps, ps.Equation.from_dsl(sde), ops, p)
    ana = ps.Equation.from_dsl("name = a\nkind = analytical\nparams = ka, ke, v, tlag, f_oral, s\nstates = depot, central\noutputs = cp\nbolus(oral) -> depot\n"
                               "infusion(iv) -> central\nlag(oral) = tlag\nfa(oral) = f_oral\nstructure = one_compartment_with_absorption\n"
                               "out(cp) = central / v ~ continuous()\n")
    want = gpu_predictions(ps, ana, ops, p)
    assert got[0] == 0.0 and got[1] == 0.0                      # nothing absorbed before the lagged dose (t = 0.5: obs sorts first)
    assert np.max(np.abs(got - want)) <= 0.03 * np.max(want)     # EM at rtol = atol = 1e-2 (sde/mod.rs:172)
