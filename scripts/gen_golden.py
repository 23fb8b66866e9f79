"""Generate the committed golden vectors under tests/golden/ from INDEPENDENT mathematics
(no oracle, no product):

  kernels.json     one-step closed-form kernels vs scipy.linalg.expm of the augmented generator
                   [[A, b], [0, 0]] (exact for linear systems with constant input)
  timelines.json   the reference's analytical unit-test fixtures (analytical/mod.rs:446-487 with the
                   parameters of each *_models.rs test) propagated event-by-event with expm
  stiff_c4.json    Michaelis-Menten + effect compartment predictions from SciPy Radau, rtol=1e-12
  normal.json      log-pdf / log-cdf / log-sf anchors from mpmath (50 digits)

Run:  python scripts/gen_golden.py     (deterministic; needs numpy, scipy, mpmath)
"""
import json
import os

import mpmath as mp
import numpy as np
from scipy.integrate import solve_ivp
from scipy.linalg import expm

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
rng = np.random.Generator(np.random.Philox(key=[20261018, 99]))


import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
from golden_math import generator, step  # noqa: E402

NPAR = {"one_compartment": 1, "one_compartment_with_absorption": 2, "two_compartments": 3, "two_compartments_with_absorption": 4,
        "three_compartments": 5, "three_compartments_with_absorption": 6}

# ---- kernels.json ---------------------------------------------------------------------------------
cases = []
for kernel, npar in NPAR.items():
    for _ in range(40):
        p = (0.05 + 2.0 * rng.random(npar)).tolist()
        n = generator(kernel, p)[0].shape[0]
        x = (100.0 * rng.random(n)).tolist()
        dt = float(rng.choice([0.25, 0.5, 1.0, 2.5, 6.0, 12.0]))
        rate = float(rng.choice([0.0, 0.0, 5.0, 50.0]))
        cases.append(dict(kernel=kernel, p=p, x=x, dt=dt, rate=rate, out=step(kernel, p, np.array(x), dt, rate).tolist()))
json.dump(cases, open(os.path.join(OUT, "kernels.json"), "w"), indent=0)

# ---- timelines.json: fixtures of analytical/mod.rs:446-487 -------------------------------------------
INFUSION_DOSING = dict(
    ops=[("bolus", 0.0, 100.0, "0"), ("infusion", 24.0, 150.0, "0", 3.0)] +
        [("missing_observation", t, "0") for t in [0, 1, 2, 4, 8, 12, 24, 25, 26, 27, 28, 32, 36]])
ORAL_INFUSION = dict(
    ops=[("bolus", 0.0, 100.0, "1"), ("infusion", 24.0, 150.0, "0", 3.0), ("bolus", 48.0, 100.0, "0")] +
        [("missing_observation", t, "0") for t in [0, 1, 2, 4, 8, 12, 24, 25, 26, 27, 28, 32, 36, 48, 49, 50, 52, 56, 60]])


def simulate(kernel, kp, v, out_state, ops):
    """Independent event loop: observations before doses at equal times; bolus to state[input];
    infusion rate feeds the kernel's input compartment; expm between consecutive breakpoints."""
    n = generator(kernel, kp)[0].shape[0]
    ev = []
    for k, op in enumerate(ops):
        rank = {"missing_observation": 0, "bolus": 1, "infusion": 2}[op[0]]
        ev.append((op[1], rank, k, op))
    ev.sort(key=lambda e: (e[0], e[1], e[2]))
    infs = [(op[1], op[1] + op[4], op[2] / op[4]) for op in ops if op[0] == "infusion"]
    x = np.zeros(n)
    preds = []
    for i, (t, rank, _, op) in enumerate(ev):
        if op[0] == "bolus":
            x[int(op[3])] += op[2]
        elif op[0] == "missing_observation":
            preds.append(x[out_state] / v)
        if i + 1 < len(ev):
            tn = ev[i + 1][0]
            bps = sorted({t, tn} | {b for s, e, _ in infs for b in (s, e) if t < b < tn})
            for a, b in zip(bps[:-1], bps[1:]):
                rate = sum(r for s, e, r in infs if a >= s and b <= e)
                x = step(kernel, kp, x, b - a, rate)
    return preds


timelines = []
for kernel, params, kp, v, out_state, fixture in [
    ("one_compartment", [0.1, 1.0], [0.1], 1.0, 0, INFUSION_DOSING),
    ("two_compartments", [0.1, 3.0, 1.0, 1.0], [0.1, 3.0, 1.0], 1.0, 0, INFUSION_DOSING),
    ("three_compartments", [0.1, 3.0, 2.0, 1.0, 0.5, 1.0], [0.1, 3.0, 2.0, 1.0, 0.5], 1.0, 0, INFUSION_DOSING),
    ("one_compartment_with_absorption", [1.0, 0.1, 1.0], [1.0, 0.1], 1.0, 1, ORAL_INFUSION),
    ("two_compartments_with_absorption", [0.1, 1.0, 3.0, 1.0, 1.0], [0.1, 1.0, 3.0, 1.0], 1.0, 1, ORAL_INFUSION),
    ("three_compartments_with_absorption", [1.0, 0.1, 3.0, 2.0, 1.0, 0.5, 1.0], [1.0, 0.1, 3.0, 2.0, 1.0, 0.5], 1.0, 1, ORAL_INFUSION),
]:
    timelines.append(dict(kernel=kernel, params=params, ops=fixture["ops"], predictions=simulate(kernel, kp, v, out_state, fixture["ops"])))
json.dump(timelines, open(os.path.join(OUT, "timelines.json"), "w"), indent=0)

# ---- stiff_c4.json: Radau rtol=1e-12 ---------------------------------------------------------------------
stiff = []
for trial in range(6):
    vmax, km, v, ke0, emax, ec50 = [float(z) for z in (10 + 50 * rng.random(), 0.2 + 4.8 * rng.random(), 15 + 45 * rng.random(),
                                                        [0.5, 5.0, 20.0, 50.0, 50.0, 35.0][trial], 50 + 100 * rng.random(), 1 + 9 * rng.random())]
    load, inf = 200.0, 150.0
    t_obs = [0.5, 1, 2, 4, 8, 12, 18, 24]
    segs = [(0.0, 6.0, 0.0), (6.0, 7.0, inf), (7.0, 12.0, 0.0), (12.0, 13.0, inf), (13.0, 18.0, 0.0), (18.0, 19.0, inf), (19.0, 24.0, 0.0)]
    x = np.array([load, 0.0])
    cp, eff = {}, {}
    for a, b, r in segs:
        def rhs(t, y, r=r):
            c = y[0] / v
            return [-vmax * c / (km + c) + r, ke0 * (c - y[1])]
        ts = [t for t in t_obs if a < t <= b]
        sol = solve_ivp(rhs, (a, b), x, method="Radau", rtol=1e-12, atol=1e-14, t_eval=ts + ([b] if (not ts or ts[-1] != b) else []))
        for k, t in enumerate(ts):
            cp[t] = sol.y[0, k] / v
            eff[t] = emax * sol.y[1, k] / (ec50 + sol.y[1, k])
        x = sol.y[:, -1]
    ops = [("bolus", 0.0, load, "load")] + [("infusion", t0, inf, "iv", 1.0) for t0 in (6.0, 12.0, 18.0)]
    preds = []
    for k, t in enumerate(t_obs):
        ops.append(("missing_observation", float(t), "cp" if k % 2 == 0 else "effect"))
        preds.append(cp[t] if k % 2 == 0 else eff[t])
    stiff.append(dict(params=[vmax, km, v, ke0, emax, ec50], ops=ops, predictions=preds))
json.dump(stiff, open(os.path.join(OUT, "stiff_c4.json"), "w"), indent=0)

# ---- normal.json -----------------------------------------------------------------------------------------------
mp.mp.dps = 50
norm = []
for obs, pred, sigma in [(0, 0, 1), (1.0, 1.0, 0.3), (2.5, 1.0, 0.5), (0.1, 4.0, 0.2), (10.0, 2.0, 0.2), (-3.0, 0.5, 1.5), (1.0, 40.0, 1.0), (40.0, 1.0, 1.0)]:
    z = (mp.mpf(obs) - mp.mpf(pred)) / mp.mpf(sigma)
    logpdf = -mp.log(2 * mp.pi) / 2 - mp.log(sigma) - z * z / 2
    cdf = mp.ncdf(z)
    sf = mp.ncdf(-z)
    norm.append(dict(obs=obs, pred=pred, sigma=sigma, logpdf=float(logpdf), logcdf=float(mp.log(cdf)), logsf=float(mp.log(sf))))
json.dump(norm, open(os.path.join(OUT, "normal.json"), "w"), indent=0)
print("wrote", sorted(os.listdir(OUT)))

# ---- dsl_features.json: a DSL model exercising the front end (array states, constants, statement-level if /
# else-if, conditional expression, intrinsics, explicit rate(), time-dependent RHS, lag, fa) against an
# independent SciPy DOP853 integration (rtol 1e-13) of the same equations written by hand --------------------
DSL_FEATURES = """
name = dsl_features
kind = ode
params = ktr, cl, v, vmax, km, f_oral
covariates = wt@linear
const wt_ref = 70.0
states = depot, transit[3], central
derived = cl_i, sat, scale
outputs = cp, lncp

bolus(oral) -> depot
infusion(iv) -> central

lag(oral) = 0.25
fa(oral) = min(max(f_oral, 0.0), 1.0)

scale = pow(wt / wt_ref, 0.75)
if (wt > 80) {
    cl_i = cl * scale * 1.1
} else if (wt > 60) {
    cl_i = cl * scale
} else {
    cl_i = cl * scale * 0.9
}
sat = if (central / v > 2.0) { vmax * (central / v - 2.0) / (km + central / v) } else { 0.0 }

dx(depot) = -ktr * depot
dx(transit[0]) = ktr * depot - ktr * transit[0]
dx(transit[1]) = ktr * (transit[0] - transit[1])
dx(transit[2]) = ktr * (transit[1] - transit[2])
dx(central) = ktr * transit[2] - (cl_i / v + sat) * central + rate(iv) * (1 - 0.1 * exp(-t / 4))

out(cp) = central / v ~ continuous()
out(lncp) = log(abs(central / v) + 1e-9) + sqrt(2 ^ 2) - 2 ~ continuous()
"""


def dsl_features_truth(p, wt0, wt1, doses, infusions, obs):
    """doses: [(t, amount)] oral; infusions: [(t, amount, dur)]; obs: [(t, outeq)]; wt linear (0, wt0) -> (48, wt1)."""
    ktr, cl, v, vmax, km, f_oral = p
    fa = min(max(f_oral, 0.0), 1.0)

    def wt(t):
        return wt0 + (wt1 - wt0) * min(t, 48.0) / 48.0

    def rhs(t, y, rate):
        w = wt(t)
        scale = (w / 70.0) ** 0.75
        cl_i = cl * scale * (1.1 if w > 80 else 1.0 if w > 60 else 0.9)
        conc = y[4] / v
        sat = vmax * (conc - 2.0) / (km + conc) if conc > 2.0 else 0.0
        return [-ktr * y[0], ktr * y[0] - ktr * y[1], ktr * (y[1] - y[2]), ktr * (y[2] - y[3]),
                ktr * y[3] - (cl_i / v + sat) * y[4] + rate * (1 - 0.1 * np.exp(-t / 4))]
    # Reference quirk restated (ode/mod.rs:343-347, 641-687; dsl/native.rs:1286): the ODE solver clock starts at
    # occasion.initial_time() computed from the UNLAGGED events and the first event is processed without
    # advancing the clock, so a lagged bolus that is the first event of its occasion lands at t0 (its lag is
    # effectively ignored); every later bolus is delayed by the lag.
    ev = [(t + (0.25 if i > 0 else 0.0), 1, ("bolus", a * fa)) for i, (t, a) in enumerate(doses)] + [(t, 0, ("obs", o)) for t, o in obs]
    bps = sorted({t for t, _, _ in ev} | {b for t, a, d in infusions for b in (t, t + d)})
    ev.sort(key=lambda e: (e[0], e[1]))
    y = np.zeros(5)
    out = []
    tcur = bps[0]
    k = 0
    for b in bps:
        if b > tcur:
            rate = sum(a / d for t, a, d in infusions if t <= tcur and b <= t + d)
            sol = solve_ivp(lambda t, yy: rhs(t, yy, rate), (tcur, b), y, method="DOP853", rtol=1e-13, atol=1e-13)
            y = sol.y[:, -1]
            tcur = b
        while k < len(ev) and ev[k][0] == b:
            kind, payload = ev[k][2]
            if kind == "obs":
                c = y[4] / v
                out.append(c if payload == "cp" else np.log(abs(c) + 1e-9) + 2.0 - 2.0)
            else:
                y[0] += payload
            k += 1
    return out


feat = []
for wt0, wt1 in [(85.0, 95.0), (64.0, 76.0), (50.0, 58.0)]:
    doses = [(0.0, 300.0), (12.0, 200.0), (24.0, 250.0)]
    infusions = [(6.0, 400.0, 2.0), (30.0, 300.0, 1.5)]
    obs = [(0.5, "cp"), (1.0, "cp"), (2.0, "lncp"), (4.0, "cp"), (6.5, "cp"), (8.0, "cp"), (12.25, "cp"), (13.0, "lncp"), (20.0, "cp"),
           (30.5, "cp"), (31.5, "cp"), (36.0, "lncp"), (47.0, "cp")]
    ops = [("covariate", "wt", 0.0, wt0), ("covariate", "wt", 48.0, wt1)]
    ops += [("bolus", t, a, "oral") for t, a in doses] + [("infusion", t, a, "iv", d) for t, a, d in infusions]
    ops += [("missing_observation", t, o) for t, o in obs]
    for p in [[1.5, 4.0, 30.0, 1.2, 1.5, 0.8], [0.6, 9.0, 55.0, 0.4, 3.0, 1.3], [3.0, 2.0, 20.0, 2.5, 0.7, 0.55]]:
        feat.append(dict(params=p, ops=ops, predictions=[float(z) for z in dsl_features_truth(p, wt0, wt1, doses, infusions, obs)]))
json.dump(dict(dsl=DSL_FEATURES, cases=feat), open(os.path.join(OUT, "dsl_features.json"), "w"), indent=0)
print("wrote dsl_features.json")

# ---- hybrid_phage.json: the stiff 6-state phage/bacteria model of the reference's long-horizon ODE regression test
# (src/simulator/equation/ode/mod.rs:1460-1517 model, :1541-1600 schedule: 1e9-unit infusions over 1.25e-3 h), first
# 13.3 h of that schedule, integrated piecewise between infusion boundaries with SciPy Radau (rtol 1e-11) --------------
PHAGE_DSL = """
name = hybrid_phage
kind = ode
params = kep, k12, k21, kdep, kcl_air, kgr, kinf, c50, klysis, burst, ksp, kdp, kn, va
const eps = 1.0e-12
const bmax = 1.0e10
states = plasma, peripheral, airway, bacc, binf, bprot
derived = phage_air, bacc_pos, binf_pos, bprot_pos, tb, inf_eff
outputs = cp, bact

infusion(iv) -> plasma

phage_air = 0.5 * (airway + sqrt(airway * airway + eps * eps))
bacc_pos = 0.5 * (bacc + sqrt(bacc * bacc + eps * eps))
binf_pos = 0.5 * (binf + sqrt(binf * binf + eps * eps))
bprot_pos = 0.5 * (bprot + sqrt(bprot * bprot + eps * eps))
tb = bacc_pos + binf_pos + bprot_pos
inf_eff = kinf * (phage_air / va) / (1.0 + (phage_air / va) / c50)

dx(plasma) = -(kep + k12 + kdep) * plasma + k21 * peripheral
dx(peripheral) = k12 * plasma - k21 * peripheral
dx(airway) = kdep * plasma - kcl_air * airway - inf_eff * bacc_pos + burst * klysis * binf_pos
dx(bacc) = kgr * bacc_pos * (1.0 - tb / bmax) - inf_eff * bacc_pos - ksp * bacc + kdp * bprot - kn * bacc
dx(binf) = inf_eff * bacc_pos - klysis * binf
dx(bprot) = ksp * bacc - kdp * bprot

init(bacc) = 3.0 * pow(10.0, 5.5)

out(cp) = plasma ~ continuous()
out(bact) = bacc ~ continuous()
"""
PHAGE_P = [20.799022436141968, 3.611151695251465, 0.20569434165954592, 3.674600839614868, 98.17452669143677, 2.072104573249817,
           1.909232258796692e-6, 427933.12072753906, 0.8622971177101135, 1.591451644897461, 4.387639760971069, 0.0917521107196808,
           1.147785520553589, 12.829959392547607]
PHAGE_INF = [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0, 5.5, 6.0, 6.5, 7.0, 7.5, 8.490833, 8.992917, 9.492917, 9.992917, 10.49292,
             10.99292, 11.49292]
PHAGE_INF3 = [12.01458, 13.01542]
PHAGE_OBS = [0.005, 0.01791667, 0.02208333, 0.03458333, 0.03833333, 0.08458333, 0.18625, 8.991667, 8.995833, 9.010417, 9.074167, 9.166667,
             12.03958, 13.01375, 13.01792, 13.1925, 13.26875]


def phage_truth():
    kep, k12, k21, kdep, kcl_air, kgr, kinf, c50, klysis, burst, ksp, kdp, kn, va = PHAGE_P
    eps, bmax = 1e-12, 1e10
    soft = lambda v: 0.5 * (v + np.sqrt(v * v + eps * eps))
    infs = [(t, 1e9, 0.00125) for t in PHAGE_INF] + [(t, 3e9, 0.00125) for t in PHAGE_INF3]

    def rhs(t, x, rate):
        pa, bp, bi, br = soft(x[2]), soft(x[3]), soft(x[4]), soft(x[5])
        tb = bp + bi + br
        cair = pa / va
        ie = kinf * cair / (1.0 + cair / c50)
        return [-(kep + k12 + kdep) * x[0] + k21 * x[1] + rate, k12 * x[0] - k21 * x[1],
                kdep * x[0] - kcl_air * x[2] - ie * bp + burst * klysis * bi,
                kgr * bp * (1.0 - tb / bmax) - ie * bp - ksp * x[3] + kdp * x[5] - kn * x[3], ie * bp - klysis * x[4], ksp * x[3] - kdp * x[5]]
    pts = sorted(set(PHAGE_OBS) | {b for t, a, d in infs for b in (t, t + d)})
    x = np.zeros(6)
    x[3] = 3.0 * 10 ** 5.5
    t = 0.0
    out = {}
    for b in pts:
        if b > t:
            rate = sum(a / d for s, a, d in infs if s <= t and b <= s + d)
            sol = solve_ivp(lambda tt, xx: rhs(tt, xx, rate), (t, b), x, method="Radau", rtol=1e-11, atol=1e-6)
            x = sol.y[:, -1]
            t = b
        if b in PHAGE_OBS:
            out[b] = (float(x[0]), float(x[3]))
    return out


truth = phage_truth()
ops = [("infusion", t, 1e9, "iv", 0.00125) for t in PHAGE_INF] + [("infusion", t, 3e9, "iv", 0.00125) for t in PHAGE_INF3]
preds = []
for t in PHAGE_OBS:
    ops.append(("missing_observation", t, "cp")); ops.append(("missing_observation", t, "bact"))
order = sorted(range(len(PHAGE_OBS)), key=lambda k: PHAGE_OBS[k])
for k in order:                         # prediction rows follow the event order: by time, cp before bact (insertion order)
    preds += [truth[PHAGE_OBS[k]][0], truth[PHAGE_OBS[k]][1]]
json.dump(dict(dsl=PHAGE_DSL, params=PHAGE_P, ops=ops, predictions=preds), open(os.path.join(OUT, "hybrid_phage.json"), "w"), indent=0)
print("wrote hybrid_phage.json")
