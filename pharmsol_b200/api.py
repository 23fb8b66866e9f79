"""Host-side mirror of pharmsol's public interface on the psi path.

Same names, argument meaning and error behaviour as the reference (citations relative to
/root/reference/src):

  Subject.builder(..).bolus/.infusion/.observation/...      data/builder.rs:84-362
  Data                                                       data/structs.rs:38
  ErrorPoly, AssayErrorModel(s)                              data/error_model.rs:17,150,786
  Equation.estimate_predictions / estimate_log_likelihood    simulator/equation/mod.rs:377-577
  ODE.with_solver / with_tolerances, OdeSolver               simulator/equation/ode/mod.rs:59-166
  log_likelihood_matrix / log_psi / psi                      simulator/likelihood/matrix.rs:52-150
  ParameterOrder                                             parameter_order.rs

Models are authored in pharmsol-dsl (text) — or with the `analytical()/ode()/sde()` builders that
mirror the `analytical!/ode!/sde!` macro declarations (pharmsol-macros/src/expand/*.rs) with DSL
expression strings in place of Rust closure bodies — and compiled to sm_100a device code.
Everything numeric runs on the GPU through the C ABI (``_lib``); there is no CPU fallback.
"""
from __future__ import annotations

import warnings
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import PharmsolError


# ---------------------------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------------------------
class Censor:
    NONE = "none"
    BLOQ = "bloq"
    ALOQ = "aloq"


class ErrorPoly:
    def __init__(self, c0, c1, c2, c3):
        self.c0, self.c1, self.c2, self.c3 = float(c0), float(c1), float(c2), float(c3)

    @classmethod
    def new(cls, c0, c1, c2, c3):
        return cls(c0, c1, c2, c3)

    def coefficients(self):
        return (self.c0, self.c1, self.c2, self.c3)


class Subject:
    """A subject = id + the builder ops that produced it (replayed through the C ABI)."""

    def __init__(self, id, ops):
        self.id = str(id)
        self.ops = list(ops)

    @staticmethod
    def builder(id):
        return SubjectBuilder(id)


class SubjectBuilder:
    def __init__(self, id):
        self.id = str(id)
        self.ops = []

    def bolus(self, time, amount, input):
        self.ops.append(("bolus", float(time), float(amount), str(input)))
        return self

    def infusion(self, time, amount, input, duration):
        self.ops.append(("infusion", float(time), float(amount), str(input), float(duration)))
        return self

    def observation(self, time, value, outeq):
        self.ops.append(("observation", float(time), float(value), str(outeq)))
        return self

    def censored_observation(self, time, value, outeq, censoring):
        self.ops.append(("censored_observation", float(time), float(value), str(outeq), censoring))
        return self

    def missing_observation(self, time, outeq):
        self.ops.append(("missing_observation", float(time), str(outeq)))
        return self

    def observation_with_error(self, time, value, outeq, errorpoly, censored=Censor.NONE):
        c = errorpoly.coefficients() if isinstance(errorpoly, ErrorPoly) else tuple(errorpoly)
        self.ops.append(("observation_with_error", float(time), float(value), str(outeq), c, censored))
        return self

    def covariate(self, name, time, value):
        self.ops.append(("covariate", str(name), float(time), float(value)))
        return self

    def repeat(self, n, delta):
        self.ops.append(("repeat", int(n), float(delta)))
        return self

    def reset(self):
        self.ops.append(("reset",))
        return self

    def fixed_covariate(self, occasion, name, fixed=True):
        """Covariate::set_fixed (data/covariate.rs:243-248) on a built occasion."""
        self.ops.append(("covariate_fixed", int(occasion), str(name), bool(fixed)))
        return self

    def build(self):
        return Subject(self.id, self.ops)


class Data:
    def __init__(self, subjects: Iterable[Subject]):
        self.subjects = list(subjects)
        self._native = None

    @classmethod
    def new(cls, subjects):
        return cls(subjects)

    def __len__(self):
        return len(self.subjects)

    def native(self):
        if self._native is None:
            self._native = _lib.NativeData(self.subjects)
        return self._native

    @classmethod
    def from_pmetrics(cls, path=None, text=None):
        """read_pmetrics (data/parser/pmetrics/mod.rs:164-239): the CSV is parsed natively (C++), ADDL/II expanded,
        occasions split at EVID=4; ``subjects`` mirrors the parsed result as builder ops."""
        return cls._from_native(_lib.NativeData.from_pmetrics(path=path, text=text))

    def expand(self, idelta, tad):
        """Data::expand (data/structs.rs:155-260): dense prediction grid of missing observations."""
        return Data._from_native(self.native().expand(idelta, tad))

    @classmethod
    def _from_native(cls, native):
        subjects = []
        cens = {0: Censor.NONE, 1: Censor.BLOQ, 2: Censor.ALOQ}
        for s in native.describe():
            ops, fixed = [], []
            for k, occ in enumerate(s["occasions"]):
                if k:
                    ops.append(("reset",))
                for name, cov in occ["covariates"].items():
                    ops += [("covariate", name, float(t), float(v)) for t, v in cov["observations"]]
                    if cov["fixed"]:
                        fixed.append(("covariate_fixed", k, name, True))
                for e in occ["events"]:
                    if e["kind"] == "bolus":
                        ops.append(("bolus", e["time"], e["amount"], e["label"]))
                    elif e["kind"] == "infusion":
                        ops.append(("infusion", e["time"], e["amount"], e["label"], e["duration"]))
                    elif e["value"] is None:
                        ops.append(("missing_observation", e["time"], e["label"]))
                    elif "errorpoly" in e:
                        ops.append(("observation_with_error", e["time"], e["value"], e["label"], tuple(e["errorpoly"]), cens[e["censoring"]]))
                    elif e["censoring"]:
                        ops.append(("censored_observation", e["time"], e["value"], e["label"], cens[e["censoring"]]))
                    else:
                        ops.append(("observation", e["time"], e["value"], e["label"]))
            subjects.append(Subject(s["id"], ops + fixed))
        data = cls(subjects)
        data._native = native
        return data


class AssayErrorModel:
    NONE, ADDITIVE, PROPORTIONAL = 0, 1, 2

    def __init__(self, kind, factor, poly):
        self.kind, self.factor, self.poly = kind, float(factor), poly

    @classmethod
    def additive(cls, poly, lambda_):
        return cls(cls.ADDITIVE, lambda_, poly)

    @classmethod
    def proportional(cls, poly, gamma):
        return cls(cls.PROPORTIONAL, gamma, poly)

    @classmethod
    def none(cls):
        return cls(cls.NONE, 0.0, ErrorPoly(0, 0, 0, 0))

    def key(self):
        return (self.kind, self.factor, self.poly.coefficients())


class AssayErrorModels:
    """Keyed by output label or dense index; bound to the equation's outputs at use time
    (error_model.rs `bind_to`)."""

    def __init__(self):
        self.models = {}

    @classmethod
    def new(cls):
        return cls()

    def add(self, outeq, model):
        if outeq in self.models:
            raise PharmsolError(15, f"error model for output `{outeq}` already exists")
        self.models[outeq] = model
        return self

    def bound(self, output_names):
        """-> dense list indexed by output equation (None where absent)."""
        dense = [None] * len(output_names)
        for key, model in self.models.items():
            if isinstance(key, (int, np.integer)):
                idx = int(key)
            else:
                key = str(key)
                if key in output_names:
                    idx = output_names.index(key)
                elif key.isdigit() and f"outeq_{key}" in output_names:
                    idx = output_names.index(f"outeq_{key}")
                elif key.isdigit():
                    idx = int(key)
                else:
                    raise PharmsolError(11, f"unknown output label `{key}` (available: {', '.join(output_names)})")
            if idx >= len(dense):
                dense.extend([None] * (idx + 1 - len(dense)))
            dense[idx] = model.key() if model.kind != AssayErrorModel.NONE else None
        return dense


# ---------------------------------------------------------------------------------------------
# equations
# ---------------------------------------------------------------------------------------------
class OdeSolver:
    """ode/mod.rs:59-84.  The reference's solvers come from diffsol: Bdf (default), Sdirk(TrBdf2 | Esdirk34),
    ExplicitRk(Tsit45) — each has a device counterpart of the same published method (psi_bdf.cuh, psi_stiff.cuh,
    psi_ode.cuh); Dopri5, Sdirk4 and the Rosenbrock method Rodas4 are this backend's own additions."""
    Dopri5 = 0
    Tsit45 = 1
    Sdirk4 = 2
    TrBdf2 = 3
    Rodas4 = 4      # Rosenbrock (linearly implicit, no Newton iteration): the stiff workhorse on the GPU
    Bdf = 5         # variable-order NDF/BDF 1-5 (the algorithm diffsol's `bdf` documents)
    Esdirk34 = 6    # ESDIRK3(4) of Jorgensen, Kristensen & Thomsen (diffsol `esdirk34`)


class CovTime:
    IntervalEnd = 0      # DSL runtime semantics (dsl/native.rs:1903-1916)
    IntervalLength = 1   # analytical! macro semantics (analytical/mod.rs:362-364), SURVEY F5


class SdeMode:
    MeanPrediction = 0   # what log_likelihood_matrix evaluates (sde/mod.rs:387-433)
    ParticleFilter = 1   # SDE::estimate_log_likelihood (sde/mod.rs:526-577, 689-736)


class EmMode:
    ReferenceAdaptive = 0
    FixedStep = 1


class EqnKind:
    ODE, Analytical, SDE = 0, 1, 2


class Prediction:
    """likelihood/prediction.rs:18-27."""

    def __init__(self, time, observation, prediction, outeq, occasion, censoring):
        self._time, self._observation, self._prediction = float(time), observation, float(prediction)
        self._outeq, self._occasion, self._censoring = int(outeq), int(occasion), censoring

    def time(self):
        return self._time

    def observation(self):
        return self._observation

    def prediction(self):
        return self._prediction

    def outeq(self):
        return self._outeq

    def occasion(self):
        return self._occasion

    def censoring(self):
        return self._censoring


class SubjectPredictions:
    """likelihood/subject.rs:19: predictions of one subject for one support point."""

    def __init__(self, predictions):
        self._predictions = list(predictions)

    def predictions(self):
        return list(self._predictions)

    def flat_predictions(self):
        return [p.prediction() for p in self._predictions]

    def flat_times(self):
        return [p.time() for p in self._predictions]

    def flat_observations(self):
        return [p.observation() for p in self._predictions]

    def __len__(self):
        return len(self._predictions)


class Equation:
    def __init__(self, source, device=None):
        self.source = source
        self._model = _lib.Model.from_dsl(source)
        self.info = self._model.info
        self.device = device
        self._pops = {}

    # -- construction ---------------------------------------------------------------------------
    @staticmethod
    def from_dsl(source, device=None):
        eq = Equation(source, device)
        cls = {0: ODE, 1: Analytical, 2: SDE}[eq._model.kind]
        eq.__class__ = cls
        eq._post_init()
        return eq

    def _post_init(self):
        pass

    def _restore(self, settings):
        pass

    @staticmethod
    def from_artifact(path, device=None):
        """load_aot_model (dsl/aot.rs:316-353): a model from a CUDA-target `.pkm` artifact; the shipped device code
        is launched as is, settings (solver, tolerances, particles ...) are the exporter's."""
        info = _lib.artifact_info(path)
        eq = Equation.__new__(Equation)
        eq._model = _lib.Model.from_artifact(path)
        eq.source, eq.info, eq.device, eq._pops = None, eq._model.info, device, {}
        eq.__class__ = {0: ODE, 1: Analytical, 2: SDE}[eq._model.kind]
        eq._post_init()
        eq._restore(info["settings"])
        eq._from_artifact = True
        return eq

    def backend(self):
        """CompiledRuntimeModel::backend (dsl/runtime.rs): which route produced the device code of this model."""
        return RuntimeBackend.CudaAot if getattr(self, "_from_artifact", False) else RuntimeBackend.Jit

    def export_artifact(self, path, solvers=()):
        """compile_module_source_to_aot's output step (dsl/aot.rs:146-300): write the `.pkm` for this model with
        its current settings; NVRTC only, no GPU needed."""
        return self._model.export_artifact(path, solvers)

    # -- introspection (equation/mod.rs:534-547) ----------------------------------------------------
    def kind(self):
        return self._model.kind

    def nstates(self):
        return self._model.nstates

    def nouteqs(self):
        return self._model.nouteqs

    def nparams(self):
        return self._model.nparams

    def parameter_names(self):
        return list(self.info["parameters"])

    def output_names(self):
        return [o["name"] for o in self.info["outputs"]]

    @property
    def cuda_source(self):
        return self._model.cuda_source

    # -- runtime ------------------------------------------------------------------------------------
    def _ctx(self):
        dev = self.device
        if dev is None:
            import os
            dev = int(os.environ.get("LOCAL_RANK", "0")) if "PHARMSOL_B200_DEVICE" not in os.environ else int(os.environ["PHARMSOL_B200_DEVICE"])
        return _lib.context(dev)

    def population(self, data: Data, error_models: AssayErrorModels | None):
        dense = error_models.bound(self.output_names()) if error_models is not None else None
        dev = self.device
        key = (id(data), repr(dense), tuple(dev) if isinstance(dev, (list, tuple)) else dev)
        pop = self._pops.get(key)
        if pop is None:
            if len(self._pops) > 8:
                self._pops.clear()
            pop = _lib.Population(self._ctx(), self._model, data.native(), dense)
            self._pops[key] = (pop, data)   # keep `data` alive so id() stays unique
            return pop
        return pop[0]

    def log_likelihood_matrix(self, data, support_points, error_models, exponentiate=False, out=None):
        pop = self.population(data, error_models)
        return _lib.log_likelihood_matrix(self._ctx(), self._model, pop, support_points, out=out, exponentiate=exponentiate)

    def predictions_matrix(self, data, support_points):
        """(nobs_total, nspp) predictions + per-subject row offsets."""
        pop = self.population(data, None)
        return _lib.predictions(self._ctx(), self._model, pop, support_points), pop.obs_offsets()

    def estimate_predictions(self, subject: Subject, parameters) -> SubjectPredictions:
        data = Data([subject])
        p = np.asarray(parameters, dtype=np.float64).reshape(1, -1)
        pred, _ = self.predictions_matrix(data, p)
        t, v, oq, oc, ce = self.population(data, None).observation_table()
        cens = {0: Censor.NONE, 1: Censor.BLOQ, 2: Censor.ALOQ}
        return SubjectPredictions(Prediction(t[k], None if v[k] != v[k] else float(v[k]), pred[k, 0], oq[k], oc[k], cens[int(ce[k])])
                                  for k in range(pred.shape[0]))

    def estimate_log_likelihood(self, subject: Subject, parameters, error_models: AssayErrorModels) -> float:
        data = Data([subject])
        p = np.asarray(parameters, dtype=np.float64).reshape(1, -1)
        return float(self.log_likelihood_matrix(data, p, error_models)[0, 0])

    def estimate_likelihood(self, subject, parameters, error_models):
        warnings.warn("Use estimate_log_likelihood() instead for better numerical stability", DeprecationWarning)
        return float(np.exp(self.estimate_log_likelihood(subject, parameters, error_models)))


class Analytical(Equation):
    def _restore(self, st):
        self._model.set_cov_time(int(st["cov_time"]))

    def with_cov_time(self, mode):
        self._model.set_cov_time(mode)
        return self


class ODE(Equation):
    def _post_init(self):
        self._solver, self._rtol, self._atol = OdeSolver.Dopri5, 1e-4, 1e-4   # RTOL/ATOL ode/mod.rs:40-41

    def _restore(self, st):
        self._solver, self._rtol, self._atol = int(st["solver"]), float(st["rtol"]), float(st["atol"])
        self._model.set_solver(self._solver, self._rtol, self._atol)

    def with_solver(self, solver):
        self._solver = int(solver)
        self._model.set_solver(self._solver, self._rtol, self._atol)
        return self

    def with_tolerances(self, rtol, atol):
        self._rtol, self._atol = float(rtol), float(atol)
        self._model.set_solver(self._solver, self._rtol, self._atol)
        return self

    def with_max_steps(self, n):
        self._model.set_max_steps(n)
        return self


class SDE(Equation):
    def _post_init(self):
        self._np = int(self.info.get("particles") or 1000)
        self._seed, self._mode, self._em, self._dt = 0x5EED, SdeMode.MeanPrediction, EmMode.ReferenceAdaptive, 0.05
        self._apply()

    def _apply(self):
        self._model.set_particles(self._np, self._seed, self._mode, self._em, self._dt)

    def _restore(self, st):
        self._np, self._seed, self._mode = int(st["nparticles"]), int(st["seed"]), int(st["sde_mode"])
        self._em, self._dt = int(st["em_mode"]), float(st["em_dt"])
        self._apply()
        self._model.set_sde_normals(int(st.get("sde_normals", 0)))

    def with_particles(self, n):
        self._np = int(n)
        self._apply()
        return self

    def with_seed(self, seed):
        self._seed = int(seed)
        self._apply()
        return self

    def with_mode(self, mode):
        self._mode = int(mode)
        self._apply()
        return self

    def with_stepper(self, em_mode, dt=0.05):
        self._em, self._dt = int(em_mode), float(dt)
        self._apply()
        return self

    def with_noise_precision(self, fp64=True):
        """Noise draws in FP64 (Box-Muller on 32-bit uniforms) instead of the default FP32 / 24-bit ones; everything else
        is FP64 either way (include/pharmsol_cuda.h PCU_SDE_NORMALS_*)."""
        self._model.set_sde_normals(1 if fp64 else 0)
        return self

    def estimate_log_likelihood(self, subject, parameters, error_models):
        """SDE::estimate_log_likelihood runs the particle filter (sde/mod.rs:689-736)."""
        saved = self._mode
        try:
            self.with_mode(SdeMode.ParticleFilter)
            return Equation.estimate_log_likelihood(self, subject, parameters, error_models)
        finally:
            self.with_mode(saved)


# ---------------------------------------------------------------------------------------------
# macro-style builders: analytical!{...} / ode!{...} / sde!{...} with DSL expression strings
# ---------------------------------------------------------------------------------------------
def _authoring(name, kind, params, covariates, states, outputs, routes, derived, body_lines, particles=None, structure=None):
    lines = [f"name = {name}", f"kind = {kind}", "params = " + ", ".join(params)]
    if covariates:
        lines.append("covariates = " + ", ".join(covariates))
    lines.append("states = " + ", ".join(states))
    if derived:
        lines.append("derived = " + ", ".join(derived))
    lines.append("outputs = " + ", ".join(outputs))
    if particles:
        lines.append(f"particles = {int(particles)}")
    for r in routes:   # "bolus(oral) -> gut"
        lines.append(r)
    if structure:
        lines.append(f"structure = {structure}")
    lines.extend(body_lines)
    return "\n".join(lines) + "\n"


def _body(derive=None, lag=None, fa=None, init=None, out=None, dx=None, noise=None):
    lines = []
    for k, v in (derive or {}).items():
        lines.append(f"{k} = {v}")
    for k, v in (lag or {}).items():
        lines.append(f"lag({k}) = {v}")
    for k, v in (fa or {}).items():
        lines.append(f"fa({k}) = {v}")
    for k, v in (dx or {}).items():
        lines.append(f"dx({k}) = {v}")
    for k, v in (noise or {}).items():
        lines.append(f"noise({k}) = {v}")
    for k, v in (init or {}).items():
        lines.append(f"init({k}) = {v}")
    for k, v in (out or {}).items():
        lines.append(f"out({k}) = {v} ~ continuous()")
    return lines


def analytical(name, params, states, outputs, routes, structure, out, covariates=(), derived=(), derive=None, lag=None, fa=None,
               init=None, device=None):
    """analytical!{ name, params, covariates, derived, states, outputs, routes, structure, derive, lag, fa, init, out }
    (pharmsol-macros/src/expand/analytical.rs:31-135)."""
    src = _authoring(name, "analytical", params, covariates, states, outputs, routes, derived,
                     _body(derive=derive, lag=lag, fa=fa, init=init, out=out), structure=structure)
    return Equation.from_dsl(src, device)


def ode(name, params, states, outputs, routes, diffeq, out, covariates=(), derived=(), derive=None, lag=None, fa=None, init=None,
        device=None):
    """ode!{ ... diffeq: {state: expr}, out: {output: expr} } (pharmsol-macros/src/expand/ode.rs:126-185)."""
    src = _authoring(name, "ode", params, covariates, states, outputs, routes, derived,
                     _body(derive=derive, lag=lag, fa=fa, init=init, out=out, dx=diffeq))
    return Equation.from_dsl(src, device)


def sde(name, params, states, outputs, routes, drift, diffusion, out, particles, covariates=(), derived=(), derive=None, lag=None,
        fa=None, init=None, device=None):
    """sde!{ ... drift, diffusion, particles } (pharmsol-macros/src/expand/sde.rs:25-117)."""
    src = _authoring(name, "sde", params, covariates, states, outputs, routes, derived,
                     _body(derive=derive, lag=lag, fa=fa, init=init, out=out, dx=drift, noise=diffusion), particles=particles)
    return Equation.from_dsl(src, device)


# ---------------------------------------------------------------------------------------------
# psi matrix (likelihood/matrix.rs)
# ---------------------------------------------------------------------------------------------
class RuntimeCompilationTarget:
    """dsl/runtime.rs:118-125.  `Jit` = NVRTC in this process; `CudaAot(path)` = export a `.pkm` and reload it (the
    reference's NativeAot round trip)."""
    Jit = "jit"

    class CudaAot:
        def __init__(self, output=None, solvers=()):
            self.output, self.solvers = output, tuple(solvers)


class RuntimeBackend:
    """dsl/mod.rs RuntimeBackend"""
    Jit = "jit"
    CudaAot = "cuda-aot"


class RuntimeArtifactFormat:
    """dsl/runtime.rs:129-133"""
    CudaAot = "cuda-aot"


def compile_module_source_to_runtime(source, target=RuntimeCompilationTarget.Jit, device=None):
    """dsl/runtime.rs:207-245"""
    eq = Equation.from_dsl(source, device)
    if isinstance(target, RuntimeCompilationTarget.CudaAot):
        import os
        import tempfile
        path = target.output or os.path.join(tempfile.mkdtemp(prefix="pharmsol_b200_"), f"model_sm_100a_{eq._model.id[:5]}.pkm")
        eq.export_artifact(path, target.solvers)
        return Equation.from_artifact(path, device)
    return eq


def compile_module_source_to_aot(source, output, solvers=(), configure=None):
    """dsl/aot.rs:146-300: DSL source -> `.pkm` on disk.  `configure(equation)` may set solver / tolerances /
    particles before the device code is generated.  Returns the output path."""
    eq = Equation.from_dsl(source)
    if configure is not None:
        configure(eq)
    return eq.export_artifact(output, solvers)


def compile_module_source_to_native_aot(source, output):
    """dsl/aot.rs:146-300 for the HOST target: DSL source -> a cdylib exporting the reference's frozen compiled-backend
    symbols (`pharmsol_dsl_api_version`, `pharmsol_dsl_model_info_json_{ptr,len}`, `pharmsol_dsl_kernel_*`), i.e. a
    `.pkm` the reference's own `load_aot_model` opens.  The emitter writes the host twin of the device code; the system
    C++ compiler builds it.  Returns the output path."""
    return Equation.from_dsl(source)._model.export_host_artifact(output)


class NativeArtifact:
    """A loaded host artifact (what `load_aot_model` does with libloading, dsl/aot.rs:316-353, 404-470): version check,
    model-info envelope, the role functions as callables `(t, states, params, covariates, routes, derived, out_len) -> out`."""
    ROLES = ("derive", "dynamics", "outputs", "init", "drift", "diffusion", "route_lag", "route_bioavailability")

    def __init__(self, path):
        import ctypes as C
        import json
        self._lib = C.CDLL(str(path))
        ver = self._lib.pharmsol_dsl_api_version
        ver.restype = C.c_uint32
        if ver() != 2:
            raise PharmsolError(15, f"artifact API version {ver()} != 2")
        ptr, ln = self._lib.pharmsol_dsl_model_info_json_ptr, self._lib.pharmsol_dsl_model_info_json_len
        ptr.restype, ln.restype = C.c_void_p, C.c_size_t
        self.envelope = json.loads(C.string_at(ptr(), ln()).decode())
        self.info = self.envelope["model"]
        self.functions = {}
        dp = C.POINTER(C.c_double)
        for role in self.ROLES:
            try:
                fn = getattr(self._lib, "pharmsol_dsl_kernel_" + role)
            except AttributeError:
                continue
            fn.restype = None
            fn.argtypes = [C.c_double, dp, dp, dp, dp, dp, dp]
            self.functions[role] = fn
        if "outputs" not in self.functions:
            raise PharmsolError(15, "artifact lacks the required symbol pharmsol_dsl_kernel_outputs")

    def call(self, role, t, states, params, covariates=(), routes=(), derived=(), out=None, out_len=None):
        import ctypes as C
        dp = C.POINTER(C.c_double)
        arr = lambda v, n=1: np.ascontiguousarray(list(v) + [0.0] * max(0, n - len(v)), dtype=np.float64)
        st, pr, cv, rt, dv = arr(states), arr(params), arr(covariates), arr(routes), arr(derived)
        if out is None:
            out = np.zeros(max(int(out_len or 1), 1))
        ptr = lambda a: a.ctypes.data_as(dp)
        self.functions[role](float(t), ptr(st), ptr(pr), ptr(cv), ptr(rt), ptr(dv), ptr(out))
        return out


def load_aot_model(path, device=None):
    """dsl/aot.rs:316-353"""
    return Equation.from_artifact(path, device)


def load_runtime_artifact(path, fmt=RuntimeArtifactFormat.CudaAot, device=None):
    """dsl/runtime.rs:247-262"""
    if fmt != RuntimeArtifactFormat.CudaAot:
        raise PharmsolError(66, f"unknown artifact format {fmt!r}")
    return Equation.from_artifact(path, device)


def read_aot_model_info(path):
    """dsl/aot.rs:303-312: metadata only (model info, settings, kernels), no device code loaded."""
    return _lib.artifact_info(path)


def log_likelihood_matrix(equation: Equation, subjects: Data, support_points, error_models: AssayErrorModels, progress: bool = False):
    """likelihood/matrix.rs:52-106.  `support_points`: rows = support points, cols = parameters in
    model order.  Returns an F-order (n_subjects, n_support_points) array of log-likelihoods; the
    first failing pair raises `PharmsolError` (matrix.rs:96-104)."""
    spp = np.asarray(support_points, dtype=np.float64)
    if progress:
        print(f"Computing log-likelihood matrix: {len(subjects)} subjects × {spp.shape[0]} support points...")
    out = equation.log_likelihood_matrix(subjects, spp, error_models)
    if progress:
        ctx = equation._ctx()
        n = len(subjects) * spp.shape[0]
        ms = ctx.last_kernel_ms
        print(f"Progress: {n}/{n} (100%) kernel {ms:.3f} ms, {n / max(ms, 1e-9) * 1e3:.3e} pairs/s")
    return out


def log_psi(equation, subjects, support_points, error_models, progress=False):
    warnings.warn("Use log_likelihood_matrix() instead", DeprecationWarning)
    return log_likelihood_matrix(equation, subjects, support_points, error_models, progress)


def psi(equation, subjects, support_points, error_models, progress=False):
    """matrix.rs:138-150: exp of the log-likelihood matrix (exponentiated on the device)."""
    warnings.warn("Use log_likelihood_matrix() instead and exponentiate if needed", DeprecationWarning)
    return equation.log_likelihood_matrix(subjects, np.asarray(support_points, dtype=np.float64), error_models, exponentiate=True)


class ResidualErrorModel:
    """data/residual_error.rs:69-139: sigma from the PREDICTION (parametric algorithms)."""
    CONSTANT, PROPORTIONAL, COMBINED, EXPONENTIAL = 1, 2, 3, 4

    def __init__(self, kind, a=0.0, b=0.0):
        self.kind, self.a, self.b = int(kind), float(a), float(b)

    @classmethod
    def constant(cls, a):
        return cls(cls.CONSTANT, a, 0.0)

    @classmethod
    def proportional(cls, b):
        return cls(cls.PROPORTIONAL, 0.0, b)

    @classmethod
    def combined(cls, a, b):
        return cls(cls.COMBINED, a, b)

    @classmethod
    def exponential(cls, sigma):
        return cls(cls.EXPONENTIAL, sigma, 0.0)


class ResidualErrorModels:
    """data/residual_error.rs:341-426: keyed by output-equation index."""

    def __init__(self):
        self.models = {}

    @classmethod
    def new(cls):
        return cls()

    def add(self, outeq, model):
        self.models[int(outeq)] = model
        return self

    def dense(self):
        n = max(self.models) + 1 if self.models else 0
        return [((m.kind, m.a, m.b) if (m := self.models.get(k)) is not None else None) for k in range(n)]


def log_likelihood_batch(equation: Equation, subjects: Data, parameters, residual_error_models: ResidualErrorModels):
    """likelihood/mod.rs:119-177: per-subject individual parameters (row i <-> subject i), prediction-based sigma.
    Returns N log-likelihoods; -inf where the simulation fails or an output has no residual model."""
    pop = equation.population(subjects, None)
    return _lib.log_likelihood_batch(equation._ctx(), equation._model, pop, parameters, residual_error_models.dense())


def read_pmetrics(path):
    """pharmsol::prelude::data::read_pmetrics."""
    return Data.from_pmetrics(path=path)


class ParameterOrder:
    """parameter_order.rs: validate an external column order once, then permute support-point
    matrices into model order."""

    def __init__(self, perm):
        self.perm = list(perm)

    @classmethod
    def with_model(cls, equation: Equation, names: Sequence[str]):
        model_names = equation.parameter_names()
        names = list(names)
        if sorted(names) != sorted(model_names):
            raise PharmsolError(15, f"parameter order {names} does not match model parameters {model_names}")
        return cls([names.index(n) for n in model_names])

    def matrix(self, support_points):
        spp = np.asarray(support_points, dtype=np.float64)
        return np.ascontiguousarray(spp[:, self.perm])


# ---------------------------------------------------------------------------------------------
# device-resident psi (inputs and outputs stay in HBM; optional column sharding over ranks)
# ---------------------------------------------------------------------------------------------
class ResidentPsi:
    """psi with everything resident in HBM: the support points are uploaded once (SoA, P x ncols) and
    every ``launch()`` is one asynchronous kernel launch on the caller's current CUDA stream writing
    the column-major block psi[:, first_col : first_col + ncols].

    With ``torch.distributed`` initialised (one process per GPU) the support-point columns are
    sharded over the ranks (``pharmsol_b200.sharding``) and ``step()`` = launch + in-place all-gather.
    torch is used for device memory, streams and the process group only.
    """

    def __init__(self, equation: Equation, data: Data, support_points, error_models: AssayErrorModels, device=None, shard=True,
                 peer_stores="auto", gather_overlap="auto", gather="auto"):
        import torch
        import torch.distributed as dist
        from .sharding import ShardedPsi
        self.torch = torch
        self.eq = equation
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        equation.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.ctx = equation._ctx()
        self.pop = equation.population(data, error_models)
        spp = np.ascontiguousarray(support_points, dtype=np.float64)
        self.nspp, self.nparams = spp.shape
        self.nsub = self.pop.nsubjects
        # How the column slabs reach every rank (N > 1):
        #   "peer"  the psi kernel stores every result into all ranks' matrices (8-byte stores, one NVLink packet each):
        #           wins when a pair costs microseconds (ODE / SDE: C2 on 8 GPUs 42.8 ms fused vs 44.3 ms with NCCL);
        #   "push"  finished column chunks are pushed to the peers by the copy engines while the next chunk computes:
        #           for closed forms, whose psi is produced at GB/s rates (C3 on 8 GPUs: 63.5 ms with peer stores,
        #           37.8 ms with a bulk NCCL all-gather whose kernel queues behind the psi CTAs);
        #   "nccl"  one in-place all_gather_into_tensor after the kernel.
        if gather == "auto" and peer_stores != "auto":
            gather = "peer" if peer_stores else "nccl"
        if gather == "auto":
            gather = "push" if equation.kind() == EqnKind.Analytical else "peer"
        peer_stores = gather in ("peer", "push")
        multi = shard and dist.is_available() and dist.is_initialized()
        if gather_overlap == "auto":
            # Phased gather (evaluate 7/8 of this rank's columns, all-gather them while the last 1/8 is evaluated;
            # sharding.ColumnPartition phases) is opt-in: on 8 GPUs C3 measured 38.3 ms phased vs 37.8 ms with one bulk
            # all-gather — at default stream priority NCCL's kernel waits behind the queued CTAs of the tail phase
            # (profiles/r01_tuning.md)
            gather_overlap = False
        gather_overlap = bool(gather_overlap) and (not peer_stores) and multi
        self.sharded = ShardedPsi(self.nsub, self.nspp, self.device, peer_stores=bool(peer_stores), tail_fraction=0.125 if gather_overlap else 0.0) if multi \
            else _SingleRank(self.nsub, self.nspp, self.device)
        if multi and peer_stores and getattr(self.sharded, "peer_ptrs", None) is None:
            gather = "nccl"          # no peer mapping available: the NCCL all-gather takes over
        self.gather_mode = gather if (multi and self.sharded.world > 1) else None
        self.ranges = [r for r in self.sharded.local_ranges]
        self.first_col = self.ranges[0][0]
        self.ncols = sum(hi - lo for lo, hi in self.ranges)
        self.columns = np.concatenate([np.arange(lo, hi) for lo, hi in self.ranges]) if self.ranges else np.zeros(0, dtype=np.int64)
        self.ld_spp = max(self.ncols, 1)
        self.spp_soa = torch.empty((self.nparams, self.ld_spp), dtype=torch.float64, device=self.device)
        if self.ncols:
            _lib.upload_support_points(self.ctx, spp[self.columns], self.spp_soa.data_ptr(), self.ld_spp, self._stream())
            torch.cuda.current_stream(self.device).synchronize()
        equation._model.compile(self.ctx)

    def _stream(self):
        """torch's current stream as a cudaStream_t; the legacy default stream (handle 0) is passed as
        cudaStreamLegacy (0x1) because NULL means "the context's own stream" in the C ABI."""
        h = self.torch.cuda.current_stream(self.device).cuda_stream
        return h if h else 1

    def _launch_phase(self, p, offset):
        lo, hi = self.ranges[p]
        if hi <= lo:
            return
        spp_ptr = self.spp_soa.data_ptr() + 8 * offset
        peers = getattr(self.sharded, "peer_ptrs", None)
        if peers is not None and self.gather_mode == "push":      # copy-engine pushes of finished chunks
            _lib.log_likelihood_matrix_push(self.ctx, self.eq._model, self.pop, spp_ptr, hi - lo, self.ld_spp, peers, self.sharded.rank, self.nsub, lo, self._stream())
            return
        if peers is not None:      # fused all-gather: results go straight into every rank's full matrix
            _lib.log_likelihood_matrix_peers(self.ctx, self.eq._model, self.pop, spp_ptr, hi - lo, self.ld_spp, peers, self.nsub, lo, self._stream())
            return
        _lib.log_likelihood_matrix_device(self.ctx, self.eq._model, self.pop, spp_ptr, hi - lo, self.ld_spp,
                                          self.sharded.local_slab(p).data_ptr(), self.nsub, lo, self._stream())

    def launch(self):
        """Asynchronous psi kernel launches for this rank's columns (one per phase; no copies, no sync)."""
        if not self.ncols:
            return
        if len(self.ranges) > 1:
            _lib.status_batch_begin(self.ctx, self._stream())
        offset = 0
        for p, (lo, hi) in enumerate(self.ranges):
            self._launch_phase(p, offset)
            offset += hi - lo

    def step(self, after_compute=None):
        """launch + all-gather of the column slabs (no-op on one rank); asynchronous.  With two phases the all-gather
        of the first runs (NCCL's own stream) while the second is evaluated.  `after_compute()` is called right after
        the last kernel launch (bench: an event that brackets the compute part)."""
        if len(self.ranges) == 1:
            self.launch()
            if after_compute:
                after_compute()
            self.sharded.gather()
            return
        if self.ncols:
            _lib.status_batch_begin(self.ctx, self._stream())
        works, offset = [], 0
        for p, (lo, hi) in enumerate(self.ranges):
            self._launch_phase(p, offset)
            offset += hi - lo
            if p == len(self.ranges) - 1 and after_compute:
                after_compute()
            works.append(self.sharded.gather_phase(p, async_op=True))
        for w in works:
            if w is not None:
                w.wait()

    def finish(self):
        """Synchronise, raise the first error over all ranks (matrix.rs:96-104), return psi as a
        (nsub, nspp) column-major device tensor view."""
        self.torch.cuda.current_stream(self.device).synchronize()
        code, pair = 0, -1
        try:
            self.ctx.collect_errors()
        except PharmsolError as e:
            code, pair = e.code, (e.pair if e.pair is not None else -1)
        code, pair = self.sharded.reduce_error(code, pair)
        if code:
            raise PharmsolError(code, f"psi evaluation failed for pair {pair}", pair)
        return self.sharded.matrix()


class _SingleRank:
    def __init__(self, nsub, nspp, device):
        import torch
        self.world, self.rank, self.nsub, self.nspp = 1, 0, nsub, nspp
        self.full = torch.empty((nspp, nsub), dtype=torch.float64, device=device)
        self.local_range = (0, nspp)
        self.local_ranges = [(0, nspp)]

    def local_slab(self, phase=-1):
        return self.full

    def gather(self):
        return self.full

    def gather_phase(self, phase, async_op=False):
        return None

    def reduce_error(self, code, pair):
        return code, pair

    def matrix(self):
        return self.full.t()
