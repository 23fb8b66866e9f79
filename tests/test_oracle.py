"""Pin the CPU oracle (test infrastructure) before trusting it as the parity checker.

Three kinds of pins (SURVEY §8c / F7 — the reference holds no literal golden vectors for the psi path):
  1. the reference's own literal anchors and unit-test fixtures, re-run through the restatement;
  2. the reference's differential tests (analytical <-> ODE twin, CL <-> rate-constant kernels);
  3. independent mathematics committed under tests/golden/ (scipy expm, mpmath, SciPy Radau;
     generator: scripts/gen_golden.py).
"""
import math

import numpy as np
import pytest

import fixtures as FX
from conftest import golden


def _scaled_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


# ---- 3. independent goldens -------------------------------------------------------------------
def test_kernels_match_expm_goldens(oracle):
    """12 closed-form kernels (analytical/*_models.rs) vs expm([[A,b],[0,0]] dt)."""
    worst = 0.0
    for c in golden("kernels"):
        out = oracle.kernel_step(c["kernel"], c["x"], c["p"], c["dt"], c["rate"])
        worst = max(worst, _scaled_err(out, c["out"]))
        # componentwise, with an absolute floor for compartments that have decayed to ~0
        assert np.allclose(out, c["out"], rtol=1e-9, atol=1e-12 * max(1.0, np.max(np.abs(c["out"]))))
    assert worst < 1e-12, worst


def test_reference_fixture_timelines_match_expm(oracle):
    """analytical/mod.rs:446-487 fixtures with each *_models.rs test's parameters."""
    for t in golden("timelines"):
        m = oracle.Model(t["kernel"])
        s = oracle.Subject([tuple(o) for o in t["ops"]])
        p = m.predictions(s, t["params"])
        assert len(p) == len(t["predictions"])
        assert np.allclose(p, t["predictions"], rtol=1e-12, atol=1e-13), t["kernel"]


def test_normal_log_densities_match_mpmath(oracle):
    for c in golden("normal"):
        assert oracle.lognormpdf(c["obs"], c["pred"], c["sigma"]) == pytest.approx(c["logpdf"], rel=1e-14, abs=1e-15)
        # BLOQ = ln Phi(z); ALOQ = ln(1 - Phi(z));  |z| > 37 takes the reference's asymptotic branch
        # (distributions.rs:60-70, 95-103) which is only a leading-order expansion
        z = (c["obs"] - c["pred"]) / c["sigma"]
        tol = 1e-12 if abs(z) < 37 else 1e-3
        # Reference quirk restated literally: the asymptotic branch is lognormpdf(obs,pred,sigma) - ln|z|,
        # i.e. the density w.r.t. obs, so it carries an extra -ln(sigma) relative to the true tail.
        quirk = -math.log(c["sigma"]) if abs(z) >= 37 else 0.0
        # erfc-based cdf loses relative accuracy only through 1 - cdf near cdf ~ 1
        if z < 5:
            assert oracle.lognormcdf(c["obs"], c["pred"], c["sigma"]) == pytest.approx(c["logcdf"] + quirk, rel=tol, abs=1e-15)
        if z > -5:
            assert oracle.lognormccdf(c["obs"], c["pred"], c["sigma"]) == pytest.approx(c["logsf"] + quirk, rel=tol, abs=1e-15)


def test_stiff_c4_oracle_matches_radau(oracle):
    m = oracle.Model("c4_mm_effect", solver="dopri5", rtol=1e-11, atol=1e-12)
    for c in golden("stiff_c4"):
        s = oracle.Subject([tuple(o) for o in c["ops"]])
        p = m.predictions(s, c["params"])
        assert np.allclose(p, c["predictions"], rtol=2e-8, atol=1e-10)


# ---- 1. literal anchors of the reference ---------------------------------------------------------
def test_lognormpdf_anchor(oracle):
    # likelihood/distributions.rs:112-118
    assert oracle.lognormpdf(0.0, 0.0, 1.0) == pytest.approx(-0.9189385332046727, abs=1e-15)


def test_cdf_at_mean_is_half(oracle):
    # likelihood/distributions.rs:141-150
    assert oracle.lognormcdf(1.0, 1.0, 0.3) == pytest.approx(math.log(0.5), abs=1e-15)
    assert oracle.lognormccdf(1.0, 1.0, 0.3) == pytest.approx(math.log(0.5), abs=1e-15)


def test_cdf_tails_finite(oracle):
    # likelihood/distributions.rs:152-181: finite at +-40 sigma through the asymptotic branch
    assert math.isfinite(oracle.lognormcdf(-40.0, 0.0, 1.0))
    assert math.isfinite(oracle.lognormccdf(40.0, 0.0, 1.0))


def test_seq_eq_accumulation_is_exactly_2p5(oracle):
    # analytical/mod.rs:492-527
    m = oracle.Model("seq_eq_accumulation")
    s = oracle.Subject([("bolus", 0.0, 0.0, "0"), ("infusion", 0.25, 1.0, "0", 0.25), ("missing_observation", 1.0, "0")])
    assert m.predictions(s, [1.0])[0] == 2.5


def test_event_ordering(oracle):
    # data/structs.rs:1148-1252: time, then Observation < Bolus < Infusion, stable
    s = oracle.Subject([("infusion", 0.0, 500.0, "1", 1.0), ("bolus", 0.0, 100.0, "1"), ("observation", 0.0, 0.0, "1")])
    assert [e[0] for e in s.events()] == ["observation", "bolus", "infusion"]
    s = oracle.Subject([("bolus", 0.0, 100.0, "1"), ("observation", 0.0, 0.0, "1"), ("observation", 1.0, 5.0, "1"),
                        ("observation", 2.0, 3.0, "1"), ("bolus", 2.0, 100.0, "1")])
    assert [(e[0], e[1]) for e in s.events()] == [("observation", 0.0), ("bolus", 0.0), ("observation", 1.0),
                                                   ("observation", 2.0), ("bolus", 2.0)]
    s = oracle.Subject([("observation", 2.0, 1.0, "1"), ("bolus", 1.0, 100.0, "1")])
    assert [e[0] for e in s.events()] == ["bolus", "observation"]


def test_covariate_interpolation(oracle):
    # data/covariate.rs:535-562
    s = oracle.Subject([("covariate", "weight", 0.0, 70.0), ("covariate", "weight", 12.0, 72.0), ("covariate", "weight", 24.0, 75.0),
                        ("covariate", "age", 0.0, 35.0), ("observation", 1.0, 1.0, "0"), ("covariate_fixed", 0, "age", True)])
    for t, v in [(0.0, 70.0), (6.0, 71.0), (12.0, 72.0), (18.0, 73.5), (24.0, 75.0), (30.0, 75.0)]:
        assert s.covariate("weight", t) == v
    for t in (0.0, 12.0, 100.0):
        assert s.covariate("age", t) == 35.0


def test_lag_and_bioavailability(oracle):
    # data/structs.rs:1270-1345 through a model: the corpus analytical case has lag 0.5 and fa 0.8:
    # an observation before the lagged dose time sees nothing; amounts are scaled by fa.
    src, twin, p, ops, _ = FX.CORPUS["analytical"]
    m = oracle.Model(twin)
    s = oracle.Subject([("bolus", 0.0, 100.0, "oral"), ("missing_observation", 0.25, "cp"), ("missing_observation", 0.5, "cp"),
                        ("missing_observation", 1.5, "cp")])
    pr = m.predictions(s, p)
    assert pr[0] == 0.0 and pr[1] == 0.0        # obs at the lagged time sorts before the bolus
    nolag = m.predictions(oracle.Subject([("bolus", 0.0, 80.0, "oral"), ("missing_observation", 1.0, "cp")]), [1.0, 0.15, 25.0, 0.0, 1.0])
    assert pr[2] == pytest.approx(nolag[0], rel=1e-14)


def test_likelihood_case_analytical_vs_ode(oracle):
    # tests/ode_optimizations.rs:1105-1184: exp(ll) of 1-cpt analytical and ODE within 1 %
    c = FX.LIKELIHOOD_CASE
    s = oracle.Subject(c["ops"])
    em = oracle.ErrorModels([c["error_model"]])
    la = oracle.Model("one_compartment").log_likelihood(s, c["params"], em)
    lo = oracle.Model("ode_one_compartment").log_likelihood(s, c["params"], em)
    assert math.exp(la) == pytest.approx(math.exp(lo), rel=1e-2)
    # closed form: x(t) = 100 e^{-0.1 t} / 50, sigma = 0.1 * obs
    ll = 0.0
    for t, o in [(1, 1.8), (2, 1.6), (4, 1.3), (8, 0.8)]:
        pred, sig = 100 * math.exp(-0.1 * t) / 50, 0.1 * o
        ll += -0.5 * math.log(2 * math.pi) - math.log(sig) - (o - pred) ** 2 / (2 * sig * sig)
    assert la == pytest.approx(ll, rel=1e-13)


# ---- 2. the reference's differential tests -----------------------------------------------------------
@pytest.mark.parametrize("kernel", list(FX.KERNEL_FIXTURES))
def test_analytical_vs_ode_twin(oracle, kernel):
    """equation/analytical/*_models.rs unit tests: analytical kernel vs handwritten ODE twin on the
    InfusionDosing / OralInfusionDosage fixtures (reference tolerance 1e-2..1e-4 with BDF at
    1e-4; here the explicit RK runs at 1e-10 so the agreement is much tighter)."""
    params, ops = FX.KERNEL_FIXTURES[kernel]
    s = oracle.Subject(ops)
    pa = oracle.Model(kernel).predictions(s, params)
    po = oracle.Model("ode_" + kernel, solver="tsit45", rtol=1e-10, atol=1e-10).predictions(s, params)
    assert np.allclose(pa, po, rtol=1e-7, atol=1e-8)
    pd = oracle.Model("ode_" + kernel, solver="dopri5", rtol=1e-10, atol=1e-10).predictions(s, params)
    assert np.allclose(pa, pd, rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("kernel,cl_params,rate_params", [
    ("one_compartment", [0.1, 1.0], [0.1, 1.0]),
    ("two_compartments", [0.1, 3.0, 1.0, 3.0], [0.1, 3.0, 1.0, 1.0]),                       # cl,q,vc,vp -> ke=.1,kcp=3,kpc=1
    ("three_compartments", [0.1, 3.0, 2.0, 1.0, 3.0, 4.0], [0.1, 3.0, 2.0, 1.0, 0.5, 1.0]),  # three_compartment_cl_models.rs
])
def test_cl_variants_equal_rate_constant_kernels(oracle, kernel, cl_params, rate_params):
    ops = FX.INFUSION_DOSING
    s = oracle.Subject(ops)
    a = oracle.Model(kernel).predictions(s, rate_params)
    b = oracle.Model(kernel + "_cl").predictions(s, cl_params)
    assert np.allclose(a, b, rtol=1e-12, atol=1e-14)


def test_ode_infusion_dose_conservation(oracle):
    # ode/mod.rs:1274-1344 (dx = rateiv only): dose conserved across short infusions / back-to-back
    m = oracle.Model("bimodal_ke", solver="tsit45", rtol=1e-6, atol=1e-6)
    p = [0.0, 1.0]
    assert m.predictions(oracle.Subject([("infusion", 0.0, 100.0, "iv", 0.1), ("observation", 0.5, 0.0, "cp")]), p)[0] == pytest.approx(100.0, rel=1e-4)
    pr = m.predictions(oracle.Subject([("infusion", 0.0, 100.0, "iv", 0.1), ("observation", 0.1, 0.0, "cp"), ("observation", 0.5, 0.0, "cp")]), p)
    assert pr[0] == pytest.approx(100.0, rel=1e-4) and pr[1] == pytest.approx(100.0, rel=1e-4)
    assert m.predictions(oracle.Subject([("infusion", 0.0, 100.0, "iv", 0.01), ("observation", 0.01, 0.0, "cp")]), p)[0] == pytest.approx(100.0, rel=1e-4)
    pr = m.predictions(oracle.Subject([("observation", 0.0, 0.0, "cp"), ("infusion", 0.5, 100.0, "iv", 0.01), ("observation", 0.52, 0.0, "cp")]), p)
    assert pr[1] == pytest.approx(100.0, rel=1e-4)
    pr = m.predictions(oracle.Subject([("infusion", 0.0, 100.0, "iv", 0.5), ("infusion", 0.5, 100.0, "iv", 0.5), ("observation", 1.0, 0.0, "cp")]), p)
    assert pr[0] == pytest.approx(200.0, rel=1e-4)


def test_c2_ode_matches_closed_form(oracle):
    from benches import workloads as W
    w = W.make("c2", nsub=3, nspp=5)
    truth = oracle.Model("c2_two_cpt_oral_analytical")
    ode = oracle.Model("c2_two_cpt_oral_ode", solver="dopri5", rtol=1e-10, atol=1e-10)
    for _, ops in w["subjects"]:
        s = oracle.Subject(ops)
        for p in w["support_points"]:
            assert np.allclose(ode.predictions(s, p), truth.predictions(s, p), rtol=1e-7, atol=1e-10)


def test_matrix_layout_and_error_propagation(oracle):
    """likelihood/matrix.rs:52-106: F-order (nsub, nspp); first error aborts."""
    from benches import workloads as W, harness as H
    w = W.make("c1", nsub=5, nspp=7)
    m, d, em = H.oracle_objects(w)
    psi = m.log_likelihood_matrix(d, w["support_points"], em)
    assert psi.shape == (5, 7) and psi.flags.f_contiguous
    one = m.log_likelihood(d.subjects[3], w["support_points"][2], em)
    assert psi[3, 2] == one
    # sigma == 0 (c0 = c1 = 0, lambda = 0) -> +inf/NaN -> NonFiniteLikelihood (prediction.rs:120-124)
    bad = oracle.ErrorModels([("additive", 0.0, (0.0, 0.0, 0.0, 0.0))])
    with pytest.raises(oracle.OracleError) as e:
        m.log_likelihood_matrix(d, w["support_points"], bad)
    assert e.value.code in (1, 2, 3)


def test_imaginary_roots_is_an_error(oracle):
    # two_compartment_models.rs:20-22 panics; the restatement raises ImaginaryRoots (code 12).
    # (ke+kcp+kpc)^2 - 4 ke kpc < 0 needs a negative rate constant
    with pytest.raises(oracle.OracleError) as e:
        oracle.kernel_step("two_compartments", [1.0, 0.0], [1.0, -3.0, 1.5], 1.0, 0.0)
    assert e.value.code == 12


def test_residual_error_model_anchors(oracle):
    # data/residual_error.rs:450-517
    c, p, cb = ("constant", 0.5, 0.0), ("proportional", 0.0, 0.1), ("combined", 0.5, 0.1)
    for pred in (0.0, 100.0, -50.0):
        assert abs(oracle.residual_sigma(c, pred) - 0.5) < 1e-10
    assert abs(oracle.residual_sigma(p, 100.0) - 10.0) < 1e-10 and abs(oracle.residual_sigma(p, -100.0) - 10.0) < 1e-10
    assert abs(oracle.residual_sigma(p, 50.0) - 5.0) < 1e-10
    assert abs(oracle.residual_sigma(cb, 0.0) - 0.5) < 1e-10 and abs(oracle.residual_sigma(cb, 100.0) - math.sqrt(100.25)) < 1e-10
    assert oracle.residual_sigma(p, 0.0) >= math.sqrt(2.220446049250313e-16) > 0.0          # sigma cutoff
    assert abs(oracle.residual_log_likelihood(("constant", 1.0, 0.0), 1.0, 0.0) - (-0.5 * (math.log(2 * math.pi) + 1.0))) < 1e-10


def test_log_likelihood_batch_semantics(oracle):
    # likelihood/mod.rs:119-177: one parameter row per subject; wrong row count is an error; a missing model is -inf
    m = oracle.Model("one_cpt_iv")
    subs = [oracle.Subject([("infusion", 0.0, 500.0, "iv", 0.5), ("observation", 1.0, 3.0, "cp"), ("missing_observation", 2.0, "cp"),
                            ("observation", 4.0, 1.5, "cp")]) for _ in range(3)]
    d = oracle.Data(subs)
    prm = [[0.3, 100.0], [0.2, 80.0], [0.5, 150.0]]
    out = oracle.log_likelihood_batch(m, d, prm, [("combined", 0.1, 0.15)])
    for i, p in enumerate(prm):
        pr = m.predictions(subs[i], p)
        want = sum(oracle.residual_log_likelihood(("combined", 0.1, 0.15), o, f) for o, f in ((3.0, pr[0]), (1.5, pr[2])))
        assert out[i] == pytest.approx(want, rel=1e-14)
    assert np.all(np.isneginf(oracle.log_likelihood_batch(m, d, prm, [None])))
    with pytest.raises(oracle.OracleError):
        oracle.log_likelihood_batch(m, d, prm[:2], [("constant", 1.0, 0.0)])


def test_assay_error_model_sigma_anchors(oracle):
    """data/error_model.rs:1186-1239: additive(ErrorPoly(1,0,0,0), lambda = 5) -> sigma = sqrt(26);
    proportional(ErrorPoly(1,0,0,0), gamma = 2) -> sigma = 2 (sigma from the OBSERVATION, here 20)."""
    m = oracle.Model("one_compartment")
    s = oracle.Subject([("bolus", 0.0, 12.0, "0"), ("observation", 1.0, 20.0, "0")])
    p = [0.2, 1.0]
    pred = m.predictions(s, p)[0]
    for em, sigma in ((("additive", 5.0, (1.0, 0.0, 0.0, 0.0)), math.sqrt(26.0)), (("proportional", 2.0, (1.0, 0.0, 0.0, 0.0)), 2.0)):
        ll = m.log_likelihood(s, p, oracle.ErrorModels([em]))
        assert ll == pytest.approx(oracle.lognormpdf(20.0, pred, sigma), rel=1e-15)
    # per-observation ErrorPoly overrides the model polynomial (error_model.rs:1056-1066)
    s2 = oracle.Subject([("bolus", 0.0, 12.0, "0"), ("observation_with_error", 1.0, 20.0, "0", (0.5, 0.1, 0.0, 0.0), "none")])
    ll = m.log_likelihood(s2, p, oracle.ErrorModels([("additive", 0.0, (1.0, 0.0, 0.0, 0.0))]))
    assert ll == pytest.approx(oracle.lognormpdf(20.0, pred, 0.5 + 0.1 * 20.0), rel=1e-15)
