"""ctypes wrapper around the CPU oracle (``oracle/liboracle.so``).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
(``pharmsol_b200``) never does.  See ``oracle/pharmsol_oracle.hpp`` for what is restated from
which reference file:line and for the parity-pinning status.

Subjects are described by plain builder-op tuples (the same tuples ``pharmsol_b200.Subject``
records), mirroring ``SubjectBuilder`` (reference ``src/data/builder.rs:84-362``)::

    ("bolus", t, amount, input) | ("infusion", t, amount, input, duration)
    ("observation", t, value, outeq) | ("missing_observation", t, outeq)
    ("censored_observation", t, value, outeq, cens) | ("observation_with_error", t, value, outeq, (c0,c1,c2,c3), cens)
    ("covariate", name, t, value) | ("repeat", n, delta) | ("reset",)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ERR_NAMES = {
    0: "OK", 1: "NonFiniteLikelihood", 2: "NegativeSigma", 3: "NonFiniteSigma", 4: "InvalidOutputEquation",
    5: "NoneErrorModel", 6: "MissingErrorModel", 7: "SolverFailure", 8: "InputOutOfRange", 9: "OuteqOutOfRange",
    10: "UnknownInputLabel", 11: "UnknownOutputLabel", 12: "ImaginaryRoots", 13: "UnsupportedInputRouteKind",
    14: "MissingCovariate", 15: "OtherError", 16: "MissingObservation",
}
CENS = {"none": 0, "bloq": 1, "aloq": 2, None: 0, 0: 0, 1: 1, 2: 2}
ERRKIND = {"none": 0, "additive": 1, "proportional": 2}
SOLVERS = {"tsit45": 0, "dopri5": 1}


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ -O3 -fopenmp)."""
    srcs = [os.path.join(_HERE, f) for f in ("capi.cpp", "models.cpp", "pharmsol_oracle.hpp", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={k: v for k, v in os.environ.items() if k not in ("CXX", "CC")})
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, d, ci, cl, cs = C.c_void_p, C.c_double, C.c_int, C.c_long, C.c_char_p
        dp, lp = C.POINTER(C.c_double), C.POINTER(C.c_long)
        sig = {
            "orc_last_error": (cs, []), "orc_max_threads": (ci, []),
            "orc_sb_new": (vp, [cs]), "orc_sb_bolus": (None, [vp, d, d, cs]), "orc_sb_infusion": (None, [vp, d, d, cs, d]),
            "orc_sb_observation": (None, [vp, d, d, cs]), "orc_sb_censored_observation": (None, [vp, d, d, cs, ci]),
            "orc_sb_missing_observation": (None, [vp, d, cs]),
            "orc_sb_observation_with_error": (None, [vp, d, d, cs, d, d, d, d, ci]),
            "orc_sb_repeat": (None, [vp, cl, d]), "orc_sb_reset": (None, [vp]), "orc_sb_covariate": (None, [vp, cs, d, d]),
            "orc_sb_build": (vp, [vp]), "orc_subject_set_covariate_fixed": (None, [vp, ci, cs, ci]),
            "orc_subject_free": (None, [vp]), "orc_subject_n_occasions": (cl, [vp]), "orc_subject_n_events": (cl, [vp, ci]),
            "orc_subject_event": (None, [vp, ci, cl, C.POINTER(ci), dp, dp]),
            "orc_covariate_interpolate": (ci, [vp, ci, cs, d, dp]),
            "orc_data_new": (vp, []), "orc_data_add": (None, [vp, vp]), "orc_data_len": (cl, [vp]), "orc_data_free": (None, [vp]),
            "orc_model_new": (vp, [cs]), "orc_model_free": (None, [vp]), "orc_model_set_solver": (None, [vp, ci, d, d]),
            "orc_model_set_particles": (None, [vp, ci]), "orc_model_kind": (ci, [vp]),
            "orc_em_new": (vp, [ci]), "orc_em_set": (None, [vp, ci, ci, d, d, d, d, d]), "orc_em_free": (None, [vp]),
            "orc_lognormpdf": (d, [d, d, d]), "orc_lognormcdf": (ci, [d, d, d, dp]), "orc_lognormccdf": (ci, [d, d, d, dp]),
            "orc_kernel_step": (ci, [cs, dp, ci, dp, ci, d, d, dp]),
            "orc_predictions": (ci, [vp, vp, dp, ci, C.c_ulonglong, dp, cl, lp, lp]),
            "orc_log_likelihood": (ci, [vp, vp, dp, ci, vp, C.c_ulonglong, dp]),
            "orc_sde_pf_log_likelihood": (ci, [vp, vp, dp, ci, vp, C.c_ulonglong, dp]),
            "orc_log_likelihood_matrix": (ci, [vp, vp, dp, cl, ci, vp, dp, ci, C.c_ulonglong, ci, lp, dp, lp]),
            "orc_log_likelihood_batch": (ci, [vp, vp, dp, cl, ci, dp, ci, dp]),
            "orc_residual_sigma": (d, [ci, d, d, d]), "orc_residual_log_likelihood": (d, [ci, d, d, d, d]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _b(s) -> bytes:
    return str(s).encode()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _check(code):
    if code != 0:
        raise OracleError(code, lib().orc_last_error().decode())


class Subject:
    """Oracle-side subject built by replaying builder ops (data/builder.rs)."""

    def __init__(self, ops, id="subject"):
        L = lib()
        b = L.orc_sb_new(_b(id))
        for op in ops:
            k = op[0]
            if k == "bolus":
                L.orc_sb_bolus(b, op[1], op[2], _b(op[3]))
            elif k == "infusion":
                L.orc_sb_infusion(b, op[1], op[2], _b(op[3]), op[4])
            elif k == "observation":
                L.orc_sb_observation(b, op[1], op[2], _b(op[3]))
            elif k == "missing_observation":
                L.orc_sb_missing_observation(b, op[1], _b(op[2]))
            elif k == "censored_observation":
                L.orc_sb_censored_observation(b, op[1], op[2], _b(op[3]), CENS[op[4]])
            elif k == "observation_with_error":
                c = op[4]
                L.orc_sb_observation_with_error(b, op[1], op[2], _b(op[3]), c[0], c[1], c[2], c[3], CENS[op[5]])
            elif k == "covariate":
                L.orc_sb_covariate(b, _b(op[1]), op[2], op[3])
            elif k == "repeat":
                L.orc_sb_repeat(b, int(op[1]), float(op[2]))
            elif k == "reset":
                L.orc_sb_reset(b)
            elif k == "covariate_fixed":
                pass  # applied after build
            else:
                raise ValueError(f"unknown subject op {op!r}")
        self.ptr = L.orc_sb_build(b)
        for op in ops:
            if op[0] == "covariate_fixed":   # ("covariate_fixed", occasion, name, fixed)
                L.orc_subject_set_covariate_fixed(self.ptr, int(op[1]), _b(op[2]), int(bool(op[3])))

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.orc_subject_free(self.ptr)
            self.ptr = None

    def events(self, occasion=0):
        L = lib()
        n = L.orc_subject_n_events(self.ptr, occasion)
        out = []
        for i in range(n):
            kind, t, a = C.c_int(), C.c_double(), C.c_double()
            L.orc_subject_event(self.ptr, occasion, i, C.byref(kind), C.byref(t), C.byref(a))
            out.append((("observation", "bolus", "infusion")[kind.value], t.value, a.value))
        return out

    def n_occasions(self):
        return lib().orc_subject_n_occasions(self.ptr)

    def covariate(self, name, t, occasion=0):
        v = C.c_double()
        rc = lib().orc_covariate_interpolate(self.ptr, occasion, _b(name), t, C.byref(v))
        if rc:
            raise KeyError(name)
        return v.value


class Data:
    def __init__(self, subjects):
        self.subjects = list(subjects)
        self.ptr = lib().orc_data_new()
        for s in self.subjects:
            lib().orc_data_add(self.ptr, s.ptr)

    def __len__(self):
        return len(self.subjects)

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.orc_data_free(self.ptr)
            self.ptr = None


class ErrorModels:
    """AssayErrorModels: list of (kind, factor, (c0,c1,c2,c3)) per output equation."""

    def __init__(self, models):
        self.ptr = lib().orc_em_new(len(models))
        for i, m in enumerate(models):
            if m is None:
                continue
            kind, factor, poly = m
            lib().orc_em_set(self.ptr, i, ERRKIND[kind], float(factor), *[float(c) for c in poly])

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.orc_em_free(self.ptr)
            self.ptr = None


class Model:
    def __init__(self, name, solver=None, rtol=None, atol=None, particles=None):
        self.ptr = lib().orc_model_new(_b(name))
        if not self.ptr:
            raise OracleError(15, lib().orc_last_error().decode())
        self.name = name
        if solver is not None or rtol is not None:
            lib().orc_model_set_solver(self.ptr, SOLVERS[solver or "tsit45"], rtol or 1e-4, atol or 1e-4)
        if particles is not None:
            lib().orc_model_set_particles(self.ptr, int(particles))

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.orc_model_free(self.ptr)
            self.ptr = None

    def predictions(self, subject, params, seed=0, return_stats=False):
        p = np.ascontiguousarray(params, dtype=np.float64)
        out = np.empty(4096, dtype=np.float64)
        n = C.c_long()
        stats = (C.c_long * 3)()
        _check(lib().orc_predictions(self.ptr, subject.ptr, _dp(p), p.size, seed, _dp(out), out.size, C.byref(n), stats))
        res = out[: n.value].copy()
        return (res, tuple(stats)) if return_stats else res

    def log_likelihood(self, subject, params, em, seed=0):
        p = np.ascontiguousarray(params, dtype=np.float64)
        v = C.c_double()
        _check(lib().orc_log_likelihood(self.ptr, subject.ptr, _dp(p), p.size, em.ptr, seed, C.byref(v)))
        return v.value

    def pf_log_likelihood(self, subject, params, em, seed=0):
        p = np.ascontiguousarray(params, dtype=np.float64)
        v = C.c_double()
        _check(lib().orc_sde_pf_log_likelihood(self.ptr, subject.ptr, _dp(p), p.size, em.ptr, seed, C.byref(v)))
        return v.value

    def log_likelihood_matrix(self, data, support_points, em, nthreads=0, seed=0, sde_mode=0, return_info=False):
        """likelihood/matrix.rs:52-106: returns an F-order (nsub, nspp) array."""
        spp = np.ascontiguousarray(support_points, dtype=np.float64)
        nspp, npar = spp.shape
        out = np.zeros((len(data), nspp), dtype=np.float64, order="F")
        pair, secs = C.c_long(-1), C.c_double()
        stats = (C.c_long * 3)()
        rc = lib().orc_log_likelihood_matrix(self.ptr, data.ptr, _dp(spp), nspp, npar, em.ptr, _dp(out), nthreads, seed,
                                             sde_mode, C.byref(pair), C.byref(secs), stats)
        if rc:
            err = OracleError(rc, lib().orc_last_error().decode())
            err.pair = pair.value
            raise err
        if return_info:
            return out, {"seconds": secs.value, "threads": nthreads or lib().orc_max_threads(),
                         "nsteps": stats[0], "nrej": stats[1], "nrhs": stats[2]}
        return out


RESID_KIND = {"constant": 1, "proportional": 2, "combined": 3, "exponential": 4, None: 0}


def residual_sigma(model, prediction):
    return lib().orc_residual_sigma(RESID_KIND[model[0]], float(model[1]), float(model[2]), float(prediction))


def residual_log_likelihood(model, observation, prediction):
    return lib().orc_residual_log_likelihood(RESID_KIND[model[0]], float(model[1]), float(model[2]), float(observation), float(prediction))


def log_likelihood_batch(model, data, parameters, residual_models):
    """likelihood/mod.rs:119-177.  residual_models: list per outeq of None | (kind, a, b)
    (constant: a; proportional: b; combined: a, b; exponential: sigma in a)."""
    prm = np.ascontiguousarray(parameters, dtype=np.float64)
    res = np.array([[RESID_KIND[m[0]], m[1], m[2]] if m else [0, 0, 0] for m in residual_models], dtype=np.float64).reshape(-1)
    out = np.empty(prm.shape[0], dtype=np.float64)
    _check(lib().orc_log_likelihood_batch(model.ptr, data.ptr, _dp(prm), prm.shape[0], prm.shape[1], _dp(res), len(residual_models), _dp(out)))
    return out


def lognormpdf(o, p, s):
    return lib().orc_lognormpdf(o, p, s)


def lognormcdf(o, p, s):
    v = C.c_double()
    _check(lib().orc_lognormcdf(o, p, s, C.byref(v)))
    return v.value


def lognormccdf(o, p, s):
    v = C.c_double()
    _check(lib().orc_lognormccdf(o, p, s, C.byref(v)))
    return v.value


def kernel_step(kernel, x, p, dt, rate=0.0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty_like(x)
    _check(lib().orc_kernel_step(_b(kernel), _dp(x), x.size, _dp(p), p.size, dt, rate, _dp(out)))
    return out
