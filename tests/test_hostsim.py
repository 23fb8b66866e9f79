"""CPU checks of the DEVICE SOURCE (csrc/device/*.cuh + the emitted model code) compiled for the host through
tests/hostsim/cuda_shim.h, against the oracle and the committed goldens.  Test infrastructure: these tests say the
per-pair logic is right before a GPU runs it; the parity tests proper are the `-m gpu` ones through the C ABI."""
import numpy as np
import pytest

import fixtures as FX
from conftest import golden


@pytest.fixture(scope="module")
def HS():
    from hostsim import HostSim
    return HostSim


def _ems(w):
    return [(1 if k == "additive" else 2, f, p) for _, (k, f, p) in w["error_models"].items()]


def _scaled(a, b, nobs):
    return float(np.max(np.abs(a - b) / (np.abs(b) + nobs)))


def test_c1_and_c3_closed_forms_match_the_oracle(HS, oracle):
    from benches import harness as H, workloads as W
    for name, bar, nobs, mode in (("c1", 1e-12, 10, None), ("c3", 1e-10, 10, "interval_end"), ("c3", 1e-10, 10, "interval_length")):
        w = W.make(name, nsub=5, nspp=40)
        if mode:
            w["oracle_model"] = "c3_three_cpt_cov_" + mode
        hs = HS(w["dsl"]).set_subjects(w["subjects"])
        psi, pred, info = hs.run(w["support_points"], _ems(w), cov_time=1 if mode == "interval_length" else 0, want_pred=True)
        om, od, oe = H.oracle_objects(w)
        ref = om.log_likelihood_matrix(od, w["support_points"], oe)
        assert info["code"] == 0 and _scaled(psi, ref, nobs) <= bar
        want = np.array([om.predictions(od.subjects[2], s) for s in w["support_points"][:6]]).T
        lo = 2 * nobs
        assert np.max(np.abs(pred[lo:lo + nobs, :6] - want) / np.maximum(np.abs(want), 1e-10)) <= bar * 10


@pytest.mark.parametrize("solver,tol,bar", [("Dopri5", 1e-10, 1e-6), ("Tsit45", 1e-10, 1e-6), ("Sdirk4", 1e-9, 1e-6), ("TrBdf2", 1e-8, 2e-5),
                                            ("Rodas4", 1e-9, 1e-6), ("Bdf", 1e-10, 1e-6), ("Esdirk34", 1e-10, 1e-6)])
def test_c2_every_solver_against_the_closed_form(HS, oracle, solver, tol, bar):
    """ode/mod.rs:59-84: all of the reference's solver names (Bdf, TrBdf2, Esdirk34, Tsit45) plus this backend's own."""
    from benches import harness as H, workloads as W
    w = W.make("c2", nsub=4, nspp=16)
    hs = HS(w["dsl"]).set_subjects(w["subjects"])
    psi, _, info = hs.run(w["support_points"], _ems(w), solver=solver, rtol=tol, atol=tol)
    om, od, oe = H.oracle_objects(dict(w, oracle_model=w["oracle_truth_model"]))
    ref = om.log_likelihood_matrix(od, w["support_points"], oe)
    assert info["code"] == 0 and _scaled(psi, ref, 12) <= bar, (solver, _scaled(psi, ref, 12))


@pytest.mark.parametrize("solver,tol,bar", [("Bdf", 1e-10, 1e-6), ("Bdf", 1e-6, 1e-4), ("Esdirk34", 1e-10, 1e-6), ("Rodas4", 1e-9, 1e-6)])
def test_stiff_model_against_radau_goldens(HS, solver, tol, bar):
    from benches import workloads as W
    hs = HS(W.model_source("c4_mm_effect"))
    for c in golden("stiff_c4"):
        hs.set_subjects([("s", [tuple(o) for o in c["ops"]])])
        _, pred, info = hs.run(np.array([c["params"]]), None, solver=solver, rtol=tol, atol=tol, want_pred=True)
        want = np.array(c["predictions"])
        assert info["code"] == 0
        assert np.max(np.abs(pred[:, 0] - want) / np.maximum(np.abs(want), 1e-6)) <= bar, (solver, c["params"])


def test_bdf_work_is_bounded_and_orders_climb(HS):
    """The multistep method must actually use its history: at rtol = 1e-8 on a smooth decay a first-order method would
    need ~1e4 steps per unit time; the variable-order scheme needs a few hundred."""
    src = "name = decay\nkind = ode\nparams = k\nstates = central\noutputs = cp\nbolus(iv) -> central\ndx(central) = -k * central\nout(cp) = central ~ continuous()\n"
    hs = HS(src).set_subjects([("s", [("bolus", 0.0, 100.0, "iv"), ("missing_observation", 10.0, "cp")])])
    _, pred, info = hs.run(np.array([[0.3]]), None, solver="Bdf", rtol=1e-8, atol=1e-8, want_pred=True)
    assert abs(pred[0, 0] - 100.0 * np.exp(-3.0)) <= 1e-5
    assert info["steps"] < 600, info


def test_esdirk34_observed_order_is_three(HS):
    """Fixed-step convergence of the advancing method on y' = -y + sin t (forced through h0 with loose tolerances is not
    available, so measure the global error at two tolerances two decades apart: error ratio ~ 100^(3/4))."""
    src = "name = forced\nkind = ode\nparams = k\nstates = y\noutputs = o\nbolus(iv) -> y\ndx(y) = -k * y + sin(t)\nout(o) = y ~ continuous()\n"
    hs = HS(src).set_subjects([("s", [("bolus", 0.0, 1.0, "iv"), ("missing_observation", 4.0, "o")])])
    exact = (lambda t: 1.5 * np.exp(-t) + 0.5 * (np.sin(t) - np.cos(t)))(4.0)
    errs = []
    for tol in (1e-5, 1e-7, 1e-9):
        _, pred, info = hs.run(np.array([[1.0]]), None, solver="Esdirk34", rtol=tol, atol=tol, want_pred=True)
        errs.append(abs(pred[0, 0] - exact))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-7


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


@pytest.mark.parametrize("kernel", ["one_compartment", "two_compartments_with_absorption", "three_compartments_cl", "three_compartments_with_absorption"])
def test_timeline_program_on_random_timelines(HS, oracle, kernel):
    """The host-built timeline program (psi_types.h EV_STEP) against the oracle's literal event walk: overlapping
    infusions, exact ties between observations / boluses / infusion boundaries, several occasions, missing observations."""
    rng = np.random.default_rng(sum(kernel.encode()) + 7)
    absorb = kernel.endswith("with_absorption")
    subjects = [(f"r{i}", _random_subject(rng, absorb, int(rng.integers(1, 4)))) for i in range(10)]
    names = FX.KERNEL_PARAMS[kernel]
    nspp = 24

    def draw(name):
        if name in ("v", "vc", "vp", "v2", "v3"):
            return rng.uniform(5.0, 80.0, nspp)
        if name in ("cl", "q", "q2", "q3"):
            return rng.uniform(0.5, 20.0, nspp)
        return rng.uniform(0.02, 2.5, nspp)
    spp = np.stack([draw(n) for n in names], axis=1)
    em = ("additive", 0.05, (0.1, 0.15, 0.0, 0.0))
    hs = HS(FX.kernel_dsl(kernel)).set_subjects(subjects)
    psi, pred, info = hs.run(spp, [(1, em[1], em[2])], want_pred=True)
    om = oracle.Model(kernel)
    od = oracle.Data([oracle.Subject(o, i) for i, o in subjects])
    ref = om.log_likelihood_matrix(od, spp, oracle.ErrorModels([em]))
    nobs = max(sum(1 for o in ops if o[0] == "observation") for _, ops in subjects)
    bar = 1e-11 if kernel.startswith("three") else 1e-12
    assert info["code"] == 0 and _scaled(psi, ref, nobs) <= bar
    row = 0
    for i, (_, ops) in enumerate(subjects):
        n = sum(1 for o in ops if o[0] in ("observation", "missing_observation"))
        for j in range(4):
            want = om.predictions(od.subjects[i], spp[j])
            got = pred[row:row + n, j]
            floor = 1e-9 * max(1.0, np.max(np.abs(want)))
            assert np.max(np.abs(got - want) / np.maximum(np.abs(want), floor)) <= 2 * bar, (kernel, i, j)
        row += n


def test_lagged_model_keeps_the_per_pair_event_walk(HS, oracle):
    """A model WITH lag cannot use the timeline program (bolus times depend on the support point): the generic cursor
    path must still agree with the oracle's lag / re-sort semantics (data/structs.rs:611-690)."""
    src, twin, p, ops, _ = FX.CORPUS["analytical_full"]
    hs = HS(src).set_subjects([("s", ops)])
    _, pred, info = hs.run(np.array([p]), None, want_pred=True)
    want = oracle.Model(twin).predictions(oracle.Subject(ops), p)
    assert info["code"] == 0 and np.max(np.abs(pred[:, 0] - want) / np.maximum(np.abs(want), 1e-10)) <= 1e-12


def test_sde_kernel_source_runs_single_threaded(HS, oracle):
    """The cooperative SDE kernel with one thread per CTA: zero diffusion reproduces the deterministic solution, the
    FP32- and FP64-noise paths agree far inside one seed-to-seed standard deviation (see tests/test_gpu_sde_parity.py)."""
    from benches import workloads as W
    w = W.make("c5", nsub=2, nspp=3, particles=96)
    hs = HS(w["dsl"]).set_subjects(w["subjects"])
    ems = _ems(w)
    kw = dict(particles=96, sde_mode=1, em_mode=1, em_dt=0.02)
    a = hs.run(w["support_points"], ems, seed=5, **kw)[0]
    b = hs.run(w["support_points"], ems, seed=5, sde_normals=1, **kw)[0]
    c = hs.run(w["support_points"], ems, seed=6, **kw)[0]
    assert np.all(np.isfinite(a)) and not np.array_equal(a, b)
    assert np.abs(a - b).max() <= 1e-3 * np.abs(a - c).max()


@pytest.mark.parametrize("key,mode", [("mean_prediction", 0), ("particle_filter", 1)])
def test_reference_adaptive_stepper_matches_the_oracle_fixture(HS, key, mode):
    """The reference's adaptive Euler-Maruyama path (sde/em.rs:134-167) of the device source on the host: normals carried
    across the 4-attempt cycles, the out-of-line infusion window, the integer-pipe clamps.  Eight (subject, support point)
    pairs of the `box64` case of tests/golden/sde_c5_oracle.json (what tests/test_gpu_sde_parity.py checks on the device
    for all 64), both likelihood modes, 128 particles, 24 seeds: seed-averaged ll within 3.5 SE pooled, 4.5 SE per pair."""
    from benches import workloads as W
    from conftest import golden
    case = next(c for c in golden("sde_c5_oracle")["cases"] if c["name"] == "box64")
    w = W.make("c5", nsub=case["nsub"], nspp=case["nspp"], particles=case["particles"])
    ns, nc, nseed = 2, 4, 24
    hs = HS(w["dsl"]).set_subjects(w["subjects"][:ns])
    ems = _ems(w)
    g = np.stack([hs.run(w["support_points"][:nc], ems, seed=41000 + s, particles=case["particles"], sde_mode=mode, em_mode=0)[0] for s in range(nseed)])
    ref = case[key]
    mean_o, var_o = np.array(ref["mean"])[:ns, :nc], np.array(ref["var"])[:ns, :nc]
    assert np.isfinite(g).all()
    se = np.sqrt(g.var(axis=0, ddof=1) / nseed + var_o / case["nseed"])
    z = (g.mean(axis=0) - mean_o) / se
    assert abs(z.mean()) * np.sqrt(z.size) <= 3.5 and np.abs(z).max() <= 4.5, z


def test_two_state_particle_filter_matches_the_oracle(HS, oracle):
    """The reference's own particle-filter fixture (tests/test_pf.rs:8-59): two states, so an attempt of the adaptive
    stepper needs 6 normals and a particle's attempts run in cycles of 2 over 3 Philox blocks (psi_sde.cuh: the carry
    pattern differs from the one-state C5 model).  Seed-averaged particle-filter ll vs the restated reference, 3.5 SE,
    at three parameter values."""
    c = FX.PF_TEST
    np_, nseed = 192, 40
    spp = np.array([[0.6], [1.0], [1.6]])
    om = oracle.Model("pf_test", particles=np_)
    od = oracle.Data([oracle.Subject(c["ops"], "a")])
    oe = oracle.ErrorModels([c["error_model"]])
    o = np.stack([om.log_likelihood_matrix(od, spp, oe, seed=900 + s, sde_mode=1)[0] for s in range(nseed)])
    hs = HS(c["dsl"]).set_subjects([("a", c["ops"])])
    em = [(1, c["error_model"][1], c["error_model"][2])]
    g = np.stack([hs.run(spp, em, seed=77000 + s, particles=np_, sde_mode=1, em_mode=0)[0][0] for s in range(nseed)])
    assert np.isfinite(o).all() and np.isfinite(g).all()
    se = np.sqrt(g.var(axis=0, ddof=1) / nseed + o.var(axis=0, ddof=1) / nseed)
    z = (g.mean(axis=0) - o.mean(axis=0)) / se
    assert np.abs(z).max() <= 3.5, (z, g.mean(axis=0), o.mean(axis=0))


def test_bdf_restart_at_an_emptied_compartment(HS):
    """Regression: pair (subject 286, column 172) of the C4 workload.  An infusion switches on at t = 12 h when the
    Michaelis-Menten compartment has emptied to ~atol: Hairer's starting step comes out at ~1e-13.  The BDF driver used
    to treat that as a collapsed step (SolverFailure on the device at BASELINE size); SciPy's threshold is 10 ulp(t)."""
    from benches import workloads as W
    w = W.make("c4", nsub=2000, nspp=10000)
    hs = HS(w["dsl"]).set_subjects([w["subjects"][286]])
    spp = w["support_points"][172:173]
    _, pred, info = hs.run(spp, None, solver="Bdf", rtol=1e-6, atol=1e-6, want_pred=True)
    _, ref, _ = hs.run(spp, None, solver="Rodas4", rtol=1e-10, atol=1e-10, want_pred=True)
    assert info["code"] == 0
    e6 = np.max(np.abs(pred - ref) / (np.abs(ref) + 1e-2))
    _, pred9, info9 = hs.run(spp, None, solver="Bdf", rtol=1e-9, atol=1e-9, want_pred=True)
    e9 = np.max(np.abs(pred9 - ref) / (np.abs(ref) + 1e-2))
    assert info9["code"] == 0 and e6 <= 1e-3 and e9 <= 1e-5 and e9 < e6        # global error ~ 100 x tolerance on this stiff pair, and converging


def test_program_longer_than_the_staged_window(HS, oracle):
    """> 96 timeline-program records: the staged copy is skipped and the records are read from the global array."""
    kernel = "two_compartments"
    ops = [("infusion", float(6 * k), 100.0, "0", 1.5) for k in range(6)] + [("observation", 0.25 * k + 0.1, 1.0 + 0.01 * k, "0") for k in range(150)]
    rng = np.random.default_rng(2)
    spp = np.column_stack([rng.uniform(0.05, 1.0, 140), rng.uniform(0.05, 1.0, 140), rng.uniform(0.05, 1.0, 140), rng.uniform(5, 80, 140)])
    em = ("additive", 0.05, (0.1, 0.15, 0.0, 0.0))
    psi, _, info = HS(FX.kernel_dsl(kernel)).set_subjects([("long", ops), ("short", ops[:8])]).run(spp, [(1, em[1], em[2])])
    ref = oracle.Model(kernel).log_likelihood_matrix(oracle.Data([oracle.Subject(ops, "long"), oracle.Subject(ops[:8], "short")]), spp, oracle.ErrorModels([em]))
    assert info["code"] == 0 and np.max(np.abs(psi - ref) / (np.abs(ref) + 150)) <= 1e-12
