"""Seeded synthetic populations for the five BASELINE.json configs and the reference's
own bench shapes (benches/common/mod.rs:117-271).

Pure data: subjects are lists of SubjectBuilder ops (consumed by both ``pharmsol_b200.Subject`` and
``oracle.Subject``), support points are numpy arrays.  Synthetic observations come from a small
vectorised fixed-step RK4 integrator written here (independent of both the product and the
oracle) at the "true" parameters, times exp(0.1 z) noise.  RNG = numpy Philox, seed 20261018.
"""
from __future__ import annotations

import os

import numpy as np

SEED = 20261018
_MODELS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pharmsol_b200", "models")


def model_source(name):
    return open(os.path.join(_MODELS, name + ".pmdsl")).read()


def _rng(seed, stream):
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def _rk4_truth(f, nstate, nsub, events, t_obs, dt=1.0 / 64):
    """Integrate dx/dt = f(t, x) for `nsub` subjects at once on a uniform grid.
    events: list of (time, kind, payload): kind 'bolus' payload (state, amount[nsub]);
    kind 'rate' payload (state, delta_rate[nsub]) (infusion start/stop).  Event and observation
    times must be multiples of dt.  Returns x at each observation time: (len(t_obs), nsub, nstate)."""
    x = np.zeros((nsub, nstate))
    rate = np.zeros((nsub, nstate))
    ev = sorted(events, key=lambda e: e[0])
    out = []
    t_end = max(list(t_obs) + [e[0] for e in ev])
    n = int(round(t_end / dt))
    obs_steps = {int(round(t / dt)): k for k, t in enumerate(t_obs)}
    ev_by_step = {}
    for e in ev:
        ev_by_step.setdefault(int(round(e[0] / dt)), []).append(e)
    res = np.zeros((len(t_obs), nsub, nstate))
    for step in range(n + 1):
        t = step * dt
        if step in obs_steps:                       # observations sort before doses at equal times
            res[obs_steps[step]] = x
        for e in ev_by_step.get(step, []):
            s, val = e[2]
            if e[1] == "bolus":
                x[:, s] += val
            else:
                rate[:, s] += val
        if step == n:
            break
        k1 = f(t, x) + rate
        k2 = f(t + dt / 2, x + dt / 2 * k1) + rate
        k3 = f(t + dt / 2, x + dt / 2 * k2) + rate
        k4 = f(t + dt, x + dt * k3) + rate
        x = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    del out
    return res


def _uniform(rng, lo, hi, n):
    return lo + (hi - lo) * rng.random(n)


# ---------------------------------------------------------------------------------------------
def make_c1(nsub=1000, nspp=1000, seed=SEED):
    """C1: analytical! one-compartment IV infusion; 1 infusion (0.5 h) + 10 observations."""
    rng = _rng(seed, 1)
    t_obs = [0.5, 1, 2, 3, 4, 6, 8, 12, 18, 24]
    amt = 500.0 * (0.8 + 0.4 * rng.random(nsub))
    ke, v = 0.3, 100.0
    f = lambda t, x: -ke * x
    ev = [(0.0, "rate", (0, amt / 0.5)), (0.5, "rate", (0, -amt / 0.5))]
    truth = _rk4_truth(f, 1, nsub, ev, t_obs)[:, :, 0] / v
    obs = truth * np.exp(0.1 * rng.standard_normal(truth.shape))
    subjects = []
    for i in range(nsub):
        ops = [("infusion", 0.0, float(amt[i]), "iv", 0.5)]
        ops += [("observation", float(t), float(obs[k, i]), "cp") for k, t in enumerate(t_obs)]
        subjects.append((f"c1-{i:05d}", ops))
    spp = np.stack([_uniform(rng, 0.05, 1.5, nspp), _uniform(rng, 30.0, 300.0, nspp)], axis=1)
    return dict(name="c1", dsl=model_source("c1_one_cpt_iv"), oracle_model="one_cpt_iv", subjects=subjects, support_points=spp,
                error_models={"cp": ("additive", 0.0, (0.1, 0.1, 0.0, 0.0))}, kind="analytical",
                desc="one-compartment IV infusion analytical, 1 infusion + 10 obs")


def make_c2(nsub=500, nspp=20000, seed=SEED):
    """C2: ode! two-compartment oral, 10 boluses q12h (100 mg) + 12 observations over 120 h."""
    rng = _rng(seed, 2)
    t_obs = [0.5, 2, 6, 10, 14, 24, 36, 48, 72, 96, 108, 120]
    ka, ke, kcp, kpc, v = 1.0, 0.15, 0.1, 0.08, 50.0
    amt = 100.0 * (0.8 + 0.4 * rng.random(nsub))

    def f(t, x):
        d = np.empty_like(x)
        d[:, 0] = -ka * x[:, 0]
        d[:, 1] = ka * x[:, 0] - (ke + kcp) * x[:, 1] + kpc * x[:, 2]
        d[:, 2] = kcp * x[:, 1] - kpc * x[:, 2]
        return d
    ev = [(12.0 * k, "bolus", (0, amt)) for k in range(10)]
    truth = _rk4_truth(f, 3, nsub, ev, t_obs)[:, :, 1] / v
    obs = truth * np.exp(0.1 * rng.standard_normal(truth.shape))
    subjects = []
    for i in range(nsub):
        ops = [("bolus", 12.0 * k, float(amt[i]), "oral") for k in range(10)]
        ops += [("observation", float(t), float(obs[k, i]), "cp") for k, t in enumerate(t_obs)]
        subjects.append((f"c2-{i:05d}", ops))
    spp = np.stack([_uniform(rng, 0.5, 2.0, nspp), _uniform(rng, 0.05, 0.3, nspp), _uniform(rng, 0.02, 0.2, nspp),
                    _uniform(rng, 0.02, 0.2, nspp), _uniform(rng, 20.0, 100.0, nspp)], axis=1)
    return dict(name="c2", dsl=model_source("c2_two_cpt_oral_ode"), oracle_model="c2_two_cpt_oral_ode",
                oracle_truth_model="c2_two_cpt_oral_analytical", subjects=subjects, support_points=spp,
                error_models={"cp": ("additive", 0.0, (0.1, 0.1, 0.0, 0.0))}, kind="ode",
                desc="two-compartment oral absorption ode, 10 doses + 12 obs")


def make_c3(nsub=10000, nspp=50000, seed=SEED):
    """C3: three_compartments_with_absorption + k10 = k10_0 (wt/70)^0.75 with a 4-point linear
    weight covariate; 4 oral boluses + 2 infusions + 10 observations over 72 h."""
    rng = _rng(seed, 3)
    t_obs = [1, 2, 4, 8, 12, 24, 30, 36, 48, 72]
    ka, k10_0, k12, k13, k21, k31, v = 1.0, 0.15, 0.3, 0.1, 0.2, 0.05, 40.0
    wt_t = [0.0, 24.0, 48.0, 72.0]
    wt0 = _uniform(rng, 50.0, 100.0, nsub)
    wt = np.stack([wt0, wt0 * (1 + 0.04 * rng.standard_normal(nsub)), wt0 * (1 + 0.06 * rng.standard_normal(nsub)),
                   wt0 * (1 + 0.08 * rng.standard_normal(nsub))], axis=1)
    oral = 100.0 * (0.8 + 0.4 * rng.random(nsub))
    iv = 200.0 * (0.8 + 0.4 * rng.random(nsub))

    def wt_at(t):
        j = min(int(t // 24.0), 2)
        w = (t - wt_t[j]) / 24.0
        return wt[:, j] * (1 - w) + wt[:, j + 1] * w

    def f(t, x):
        k10 = k10_0 * (wt_at(min(t, 72.0)) / 70.0) ** 0.75
        d = np.empty_like(x)
        d[:, 0] = -ka * x[:, 0]
        d[:, 1] = ka * x[:, 0] - (k10 + k12 + k13) * x[:, 1] + k21 * x[:, 2] + k31 * x[:, 3]
        d[:, 2] = k12 * x[:, 1] - k21 * x[:, 2]
        d[:, 3] = k13 * x[:, 1] - k31 * x[:, 3]
        return d
    oral_t = [0.0, 12.0, 24.0, 36.0]
    inf_t = [(6.0, 1.0), (30.0, 2.0)]
    ev = [(t, "bolus", (0, oral)) for t in oral_t]
    for t0, dur in inf_t:
        ev += [(t0, "rate", (1, iv / dur)), (t0 + dur, "rate", (1, -iv / dur))]
    truth = _rk4_truth(f, 4, nsub, ev, t_obs, dt=1.0 / 32)[:, :, 1] / v
    obs = truth * np.exp(0.1 * rng.standard_normal(truth.shape))
    subjects = []
    for i in range(nsub):
        ops = [("covariate", "wt", wt_t[k], float(wt[i, k])) for k in range(4)]
        ops += [("bolus", t, float(oral[i]), "oral") for t in oral_t]
        ops += [("infusion", t0, float(iv[i]), "iv", dur) for t0, dur in inf_t]
        ops += [("observation", float(t), float(obs[k, i]), "cp") for k, t in enumerate(t_obs)]
        subjects.append((f"c3-{i:05d}", ops))
    spp = np.stack([_uniform(rng, 0.3, 3.0, nspp), _uniform(rng, 0.02, 0.5, nspp), _uniform(rng, 0.05, 2.0, nspp),
                    _uniform(rng, 0.01, 1.0, nspp), _uniform(rng, 0.05, 2.0, nspp), _uniform(rng, 0.005, 0.5, nspp),
                    _uniform(rng, 10.0, 100.0, nspp)], axis=1)
    return dict(name="c3", dsl=model_source("c3_three_cpt_cov"), oracle_model="c3_three_cpt_cov_interval_end", subjects=subjects,
                support_points=spp, error_models={"cp": ("additive", 0.0, (0.1, 0.1, 0.0, 0.0))}, kind="analytical",
                desc="three-compartment IV+oral analytical, time-varying weight, 4 boluses + 2 infusions + 10 obs")


def make_c4(nsub=2000, nspp=10000, seed=SEED):
    """C4: Michaelis-Menten elimination + effect compartment (stiff through ke0 up to 50 /h and
    small km); 1 loading bolus + 3 infusions + 8 observations (alternating cp / effect)."""
    rng = _rng(seed, 4)
    t_obs = [0.5, 1, 2, 4, 8, 12, 18, 24]
    vmax, km, v, ke0, emax, ec50 = 30.0, 2.0, 30.0, 5.0, 100.0, 3.0
    load = 200.0 * (0.8 + 0.4 * rng.random(nsub))
    inf = 150.0 * (0.8 + 0.4 * rng.random(nsub))

    def f(t, x):
        conc = x[:, 0] / v
        d = np.empty_like(x)
        d[:, 0] = -vmax * conc / (km + conc)
        d[:, 1] = ke0 * (conc - x[:, 1])
        return d
    ev = [(0.0, "bolus", (0, load))]
    for t0 in (6.0, 12.0, 18.0):
        ev += [(t0, "rate", (0, inf / 1.0)), (t0 + 1.0, "rate", (0, -inf / 1.0))]
    xs = _rk4_truth(f, 2, nsub, ev, t_obs, dt=1.0 / 128)
    cp = xs[:, :, 0] / v
    eff = emax * xs[:, :, 1] / (ec50 + xs[:, :, 1])
    noise = np.exp(0.1 * rng.standard_normal(cp.shape))
    subjects = []
    for i in range(nsub):
        ops = [("bolus", 0.0, float(load[i]), "load")]
        ops += [("infusion", t0, float(inf[i]), "iv", 1.0) for t0 in (6.0, 12.0, 18.0)]
        for k, t in enumerate(t_obs):
            if k % 2 == 0:
                ops.append(("observation", float(t), float(cp[k, i] * noise[k, i]), "cp"))
            else:
                ops.append(("observation", float(t), float(eff[k, i] * noise[k, i]), "effect"))
        subjects.append((f"c4-{i:05d}", ops))
    spp = np.stack([_uniform(rng, 10.0, 60.0, nspp), _uniform(rng, 0.2, 5.0, nspp), _uniform(rng, 15.0, 60.0, nspp),
                    _uniform(rng, 0.5, 50.0, nspp), _uniform(rng, 50.0, 150.0, nspp), _uniform(rng, 1.0, 10.0, nspp)], axis=1)
    return dict(name="c4", dsl=model_source("c4_mm_effect"), oracle_model="c4_mm_effect", subjects=subjects, support_points=spp,
                error_models={"cp": ("additive", 0.0, (0.1, 0.1, 0.0, 0.0)), "effect": ("additive", 0.0, (0.5, 0.1, 0.0, 0.0))},
                kind="ode", desc="stiff Michaelis-Menten + effect compartment (DSL -> NVRTC), 1 bolus + 3 infusions + 8 obs")


def make_c5(nsub=200, nspp=5000, seed=SEED, particles=1000):
    """C5: sde! one-compartment with additive diffusion; 1 infusion (1 h) + 8 observations over 12 h."""
    rng = _rng(seed, 5)
    t_obs = [0.5, 1, 2, 3, 4, 6, 8, 12]
    ke, v = 0.4, 100.0
    amt = 500.0 * (0.8 + 0.4 * rng.random(nsub))
    f = lambda t, x: -ke * x
    ev = [(0.0, "rate", (0, amt / 1.0)), (1.0, "rate", (0, -amt / 1.0))]
    truth = _rk4_truth(f, 1, nsub, ev, t_obs)[:, :, 0] / v
    obs = truth * np.exp(0.1 * rng.standard_normal(truth.shape))
    subjects = []
    for i in range(nsub):
        ops = [("infusion", 0.0, float(amt[i]), "iv", 1.0)]
        ops += [("observation", float(t), float(obs[k, i]), "cp") for k, t in enumerate(t_obs)]
        subjects.append((f"c5-{i:05d}", ops))
    spp = np.stack([_uniform(rng, 0.1, 1.0, nspp), _uniform(rng, 0.01, 0.5, nspp), _uniform(rng, 50.0, 200.0, nspp)], axis=1)
    return dict(name="c5", dsl=model_source("c5_one_cpt_sde"), oracle_model="c5_one_cpt_sde", subjects=subjects, support_points=spp,
                error_models={"cp": ("additive", 0.0, (0.1, 0.1, 0.0, 0.0))}, kind="sde", particles=particles,
                desc=f"one-compartment sde, {particles} particles, 1 infusion + 8 obs")


# ---- the reference's own criterion shapes (benches/common/mod.rs) ---------------------------------
SHORT_TIMES = [0.25, 0.5, 1.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0]
SHORT_OBS = [0.50, 0.90, 1.60, 2.40, 2.10, 1.50, 1.05, 0.72, 0.48]
REPEAT_TIMES = [0.5, 2.0, 6.0, 10.0, 14.0, 24.0, 36.0, 48.0, 60.0, 72.0, 84.0, 96.0, 108.0, 120.0]
REPEAT_OBS = [1.80, 1.45, 1.10, 0.90, 1.30, 1.60, 1.55, 1.50, 1.48, 1.45, 1.43, 1.42, 1.41, 0.95]


def reference_bench(workload="repeat", nsub=32, nspp=64):
    """matrix_data / support_points of benches/common/mod.rs:181-254 (32 x 64 by default)."""
    subjects = []
    for i in range(nsub):
        off = i * 0.01
        if workload == "short":
            ops = [("bolus", 0.0, 100.0, "po")] + [("observation", t, y + off, "plasma") for t, y in zip(SHORT_TIMES, SHORT_OBS)]
        else:
            ops = [("bolus", 12.0 * d, 100.0, "iv") for d in range(10)]
            ops += [("observation", t, y + off, "plasma") for t, y in zip(REPEAT_TIMES, REPEAT_OBS)]
        subjects.append((f"{workload}-{i:03d}", ops))
    base = [1.0, 0.2, 50.0] if workload == "short" else [0.10, 0.05, 0.04, 50.0]
    spp = np.array([[p + r * 0.001 * max(abs(p), 1e-3) for p in base] for r in range(nspp)])
    return dict(name=f"ref-{workload}", subjects=subjects, support_points=spp,
                error_models={"plasma": ("additive", 0.0, (0.1, 0.1, 0.0, 0.0))})


MAKERS = {"c1": make_c1, "c2": make_c2, "c3": make_c3, "c4": make_c4, "c5": make_c5}


def make(name, **kw):
    return MAKERS[name](**kw)
