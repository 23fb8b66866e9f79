"""Reference fixture INPUTS restated as data.  Outputs always come from the
oracle / committed goldens, never from here.

Sources (relative to the reference tree):
  tests/support/runtime_corpus.rs:22-148   DSL sources of the 5 corpus cases
  tests/support/runtime_corpus.rs:200-372  support points + subjects
  src/simulator/equation/analytical/mod.rs:446-487   InfusionDosing / OralInfusionDosage
  tests/ode_optimizations.rs:1105-1161     likelihood case
  tests/test_pf.rs:8-59                    particle-filter fixture
"""

ODE_SOURCE = """
name = one_cmt_oral_iv
kind = ode

params = ka, cl, v, tlag, f_oral
covariates = wt@linear
states = depot, central
derived = cl_i, ke
outputs = cp

bolus(oral) -> depot
infusion(iv) -> central

lag(oral) = tlag
fa(oral) = f_oral

cl_i = cl * pow(wt / 70.0, 0.75)
ke = cl_i / v

dx(depot) = -ka * depot
dx(central) = ka * depot - ke * central

out(cp) = central / v ~ continuous()
"""

ODE_FULL_SOURCE = """
name = ode_full_feature_parity
kind = ode

params = ka, ke, kcp, kpc, v, tlag, f_oral, base_depot, base_central, base_peripheral
covariates = wt@linear, renal@linear
derived = adjusted_ke, adjusted_kcp, adjusted_v
states = depot, central, peripheral
outputs = cp

bolus(oral) -> depot
bolus(load) -> central
infusion(iv) -> central

lag(oral) = tlag * sqrt(wt / 70.0) * pow(90.0 / renal, 0.1)
fa(oral) = min(max(f_oral * pow(renal / 90.0, 0.1), 0.0), 1.0)

adjusted_ke = ke * pow(wt / 70.0, 0.75) * pow(renal / 90.0, 0.25)
adjusted_kcp = kcp * pow(wt / 70.0, 0.25)
adjusted_v = v * (wt / 70.0) * (1.0 + 0.001 * (renal - 90.0))

dx(depot) = -ka * depot
dx(central) = ka * depot - (adjusted_ke + adjusted_kcp) * central + kpc * peripheral
dx(peripheral) = adjusted_kcp * central - kpc * peripheral

init(depot) = base_depot + 0.05 * wt
init(central) = base_central + 0.1 * renal
init(peripheral) = base_peripheral + 0.02 * wt

out(cp) = central / adjusted_v ~ continuous()
"""

ANALYTICAL_SOURCE = """
name = one_cmt_abs
kind = analytical

params = ka, ke, v, tlag, f_oral
states = depot, central
outputs = cp

bolus(oral) -> depot

lag(oral) = tlag
fa(oral) = f_oral

structure = one_compartment_with_absorption

out(cp) = central / v ~ continuous()
"""

ANALYTICAL_FULL_SOURCE = """
name = analytical_full_feature_parity
kind = analytical

params = ka, ke, v, tlag, f_oral, base_gut, base_central
covariates = wt@linear, renal@linear
derived = adjusted_v
states = gut, central
outputs = cp

bolus(oral) -> gut
bolus(load) -> central
infusion(iv) -> central

lag(oral) = tlag * sqrt(wt / 70.0) * pow(90.0 / renal, 0.1)
fa(oral) = min(max(f_oral * pow(renal / 90.0, 0.1), 0.0), 1.0)

adjusted_v = v * (wt / 70.0) * (1.0 + 0.001 * (renal - 90.0))

structure = one_compartment_with_absorption

init(gut) = base_gut + 0.03 * wt
init(central) = base_central + 0.08 * renal

out(cp) = central / adjusted_v ~ continuous()
"""

SDE_SOURCE = """
name = vanco_sde
kind = sde

params = ka, ke0, kcp, kpc, vol, ske
covariates = wt@locf
states = depot, central, peripheral, ke_latent
particles = 16
outputs = cp

bolus(oral) -> depot

init(ke_latent) = ke0

dx(depot) = -ka * depot
dx(central) = ka * depot - (ke_latent + kcp) * central + kpc * peripheral
dx(peripheral) = kcp * central - kpc * peripheral
dx(ke_latent) = -ke_latent + ke0

noise(ke_latent) = ske

out(cp) = central / (vol * wt) ~ continuous()
"""


def _missing(times, out="cp"):
    return [("missing_observation", float(t), out) for t in times]


_FULL_COV = [("covariate", "wt", 0.0, 68.0), ("covariate", "wt", 8.0, 74.0),
             ("covariate", "renal", 0.0, 95.0), ("covariate", "renal", 8.0, 72.0)]

# name -> (dsl source, oracle handwritten twin, support point, subject ops, reference tolerance)
CORPUS = {
    "ode": (ODE_SOURCE, "corpus_ode", [1.2, 5.0, 40.0, 0.5, 0.8],
            [("covariate", "wt", 0.0, 70.0), ("bolus", 0.0, 120.0, "oral"), ("infusion", 6.0, 60.0, "iv", 2.0)]
            + _missing([0.5, 1.0, 2.0, 6.0, 7.0, 9.0]), 1e-4),
    "ode_full": (ODE_FULL_SOURCE, "corpus_ode_full", [1.1, 0.18, 0.07, 0.04, 35.0, 0.6, 0.85, 4.0, 18.0, 9.0],
                 [("bolus", 0.0, 80.0, "load"), ("bolus", 1.0, 120.0, "oral"), ("infusion", 6.0, 150.0, "iv", 2.5)]
                 + _missing([0.25, 0.75, 1.5, 3.0, 6.5, 7.0, 8.0, 12.0]) + _FULL_COV, 1e-4),
    "analytical": (ANALYTICAL_SOURCE, "corpus_analytical", [1.0, 0.15, 25.0, 0.5, 0.8],
                   [("bolus", 0.0, 100.0, "oral")] + _missing([0.5, 1.0, 2.0, 4.0]), 1e-8),
    "analytical_full": (ANALYTICAL_FULL_SOURCE, "corpus_analytical_full", [1.0, 0.16, 32.0, 0.5, 0.8, 3.0, 14.0],
                        [("bolus", 0.0, 60.0, "load"), ("bolus", 1.0, 100.0, "oral"), ("infusion", 6.0, 140.0, "iv", 2.0)]
                        + _missing([0.25, 0.75, 1.5, 3.0, 6.5, 7.0, 8.0, 12.0]) + _FULL_COV, 1e-8),
    "sde": (SDE_SOURCE, "corpus_sde", [1.1, 0.2, 0.12, 0.08, 15.0, 0.0],
            [("covariate", "wt", 0.0, 70.0), ("bolus", 0.0, 80.0, "oral")] + _missing([0.5, 1.0, 2.0, 4.0]), 1e-4),
}

# analytical/mod.rs:446-487 — numeric labels ("0", "1") resolve through the input_<n>/outeq_<n> aliases
INFUSION_DOSING = [("bolus", 0.0, 100.0, "0"), ("infusion", 24.0, 150.0, "0", 3.0)] + \
    _missing([0, 1, 2, 4, 8, 12, 24, 25, 26, 27, 28, 32, 36], "0")
ORAL_INFUSION_DOSAGE = [("bolus", 0.0, 100.0, "1"), ("infusion", 24.0, 150.0, "0", 3.0), ("bolus", 48.0, 100.0, "0")] + \
    _missing([0, 1, 2, 4, 8, 12, 24, 25, 26, 27, 28, 32, 36, 48, 49, 50, 52, 56, 60], "0")

# kernel -> (analytical params incl. v, CL-variant name, CL params, ODE twin, fixture, nstates)
KERNEL_FIXTURES = {
    "one_compartment": ([0.1, 1.0], INFUSION_DOSING),
    "two_compartments": ([0.1, 3.0, 1.0, 1.0], INFUSION_DOSING),
    "three_compartments": ([0.1, 3.0, 2.0, 1.0, 0.5, 1.0], INFUSION_DOSING),
    "one_compartment_with_absorption": ([1.0, 0.1, 1.0], ORAL_INFUSION_DOSAGE),
    "two_compartments_with_absorption": ([0.1, 1.0, 3.0, 1.0, 1.0], ORAL_INFUSION_DOSAGE),
    "three_compartments_with_absorption": ([1.0, 0.1, 3.0, 2.0, 1.0, 0.5, 1.0], ORAL_INFUSION_DOSAGE),
}

# parameter names of each built-in structure, model order (pharmsol-dsl analysis.rs:240-255) + output volume
KERNEL_PARAMS = {
    "one_compartment": ["ke", "v"],
    "one_compartment_with_absorption": ["ka", "ke", "v"],
    "two_compartments": ["ke", "kcp", "kpc", "v"],
    "two_compartments_with_absorption": ["ke", "ka", "kcp", "kpc", "v"],
    "three_compartments": ["k10", "k12", "k13", "k21", "k31", "v"],
    "three_compartments_with_absorption": ["ka", "k10", "k12", "k13", "k21", "k31", "v"],
    "one_compartment_cl": ["cl", "v"],
    "one_compartment_cl_with_absorption": ["ka", "cl", "v"],
    "two_compartments_cl": ["cl", "q", "vc", "vp"],
    "two_compartments_cl_with_absorption": ["ka", "cl", "q", "vc", "vp"],
    "three_compartments_cl": ["cl", "q2", "q3", "vc", "v2", "v3"],
    "three_compartments_cl_with_absorption": ["ka", "cl", "q2", "q3", "vc", "v2", "v3"],
}
KERNEL_STATES = {1: ["central"], 2: ["central", "peripheral"], 3: ["central", "p1", "p2"]}


def kernel_dsl(kernel):
    """DSL twin of the unit-test models in equation/analytical/*_models.rs: output = central / V,
    bolus input 0 -> state 0, (absorption) bolus input 1 -> state 1, infusion input 0."""
    params = KERNEL_PARAMS[kernel]
    ncpt = 1 if kernel.startswith("one") else 2 if kernel.startswith("two") else 3
    absorb = kernel.endswith("with_absorption")
    states = (["gut"] if absorb else []) + KERNEL_STATES[ncpt]
    vol = "v" if "v" in params else "vc"
    lines = [f"name = fx_{kernel}", "kind = analytical", "params = " + ", ".join(params), "states = " + ", ".join(states),
             "outputs = outeq_0"]
    # numeric data labels "0"/"1" resolve through the input_<n> aliases (metadata.rs:248-275)
    lines.append(f"bolus(input_0) -> {states[0]}")
    if absorb:
        lines.append("bolus(input_1) -> central")
    lines.append("infusion(input_0) -> central")
    lines.append(f"structure = {kernel}")
    lines.append(f"out(outeq_0) = central / {vol} ~ continuous()")
    return "\n".join(lines) + "\n"


# tests/ode_optimizations.rs:1105-1161
LIKELIHOOD_CASE = dict(
    ops=[("bolus", 0.0, 100.0, "0"), ("observation", 1.0, 1.8, "0"), ("observation", 2.0, 1.6, "0"),
         ("observation", 4.0, 1.3, "0"), ("observation", 8.0, 0.8, "0")],
    params=[0.1, 50.0], error_model=("additive", 0.0, (0.0, 0.1, 0.0, 0.0)))

# tests/test_pf.rs:8-59
PF_TEST = dict(
    ops=[("bolus", 0.0, 20.0, "dose"), ("observation", 0.2, 16.6434, "cp"), ("observation", 0.4, 14.3233, "cp"),
         ("observation", 0.6, 9.8468, "cp"), ("observation", 0.8, 9.4177, "cp"), ("observation", 1.0, 7.5170, "cp")],
    params=[1.0], error_model=("additive", 0.0, (0.5, 0.0, 0.0, 0.0)),
    dsl="""
name = pf_test
kind = sde
params = ke0
states = x0, x1
outputs = cp
particles = 10000
bolus(dose) -> x0
init(x1) = 1.0
dx(x0) = -x0 * x1
dx(x1) = -x1 + ke0
noise(x0) = 1.0
noise(x1) = 0.01
out(cp) = x0 ~ continuous()
""")
