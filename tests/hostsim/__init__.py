"""tests/hostsim — TEST INFRASTRUCTURE ONLY: the device headers + one emitted model translation unit compiled for the
HOST through `cuda_shim.h`, driven one "thread" at a time.

Why it exists: the authoring container has no GPU, and the kernels' per-pair logic (event walk, closed forms, the seven
ODE solvers, likelihood epilogue) is ordinary C++ once the CUDA qualifiers are defined away.  Running that same source on
the CPU against the oracle catches logic errors before GPU minutes are spent.  It is NOT a product path and NOT a CPU
fallback: nothing under `pharmsol_b200/` imports, includes or links it; the shipped library fails loudly without a
device; parity claims rest on the `-m gpu` tests only.  SDE models (cooperative CTA code) are not simulated.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pharmsol_b200", "csrc")
# HOSTSIM_SANITIZE=1: build the host double with AddressSanitizer + UBSan (own build directory) and run it under
#   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_hostsim.py
# — the bounds / UB check of the device source that compute-sanitizer would give on a GPU (it is closed on this pool).
SANITIZE = bool(os.environ.get("HOSTSIM_SANITIZE"))
BUILD = os.path.join(HERE, "_build_san" if SANITIZE else "_build")
SAN_FLAGS = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"] if SANITIZE else []
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
DSLC = os.path.join(ROOT, "pharmsol_b200", "_build", "dslc")


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"command failed: {' '.join(cmd)}\n{r.stdout}")
    return r.stdout


def _newer(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def _common_objects():
    os.makedirs(BUILD, exist_ok=True)
    host = os.path.join(CSRC, "host")
    hdrs = [os.path.join(host, h) for h in ("data.hpp", "dsl.hpp", "launch_geometry.hpp")] + [os.path.join(CSRC, "device", "psi_types.h")]
    objs = []
    for src in (os.path.join(host, "data.cpp"), os.path.join(host, "dsl_parse.cpp"), os.path.join(host, "dsl_emit.cpp"), os.path.join(HERE, "hostsim_api.cpp")):
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        if _newer(obj, [src] + hdrs):
            _run([GXX, "-std=c++17", "-O1", "-fPIC"] + SAN_FLAGS + ["-I", host, "-I", os.path.join(CSRC, "device"), "-c", src, "-o", obj])
        objs.append(obj)
    return objs


def build_module(dsl_source: str) -> str:
    """DSL source -> emitted CUDA-C translation unit (the product's own emitter, via dslc) -> host .so."""
    from pharmsol_b200 import build as B
    B.build_dslc()
    os.makedirs(BUILD, exist_ok=True)
    dev = os.path.join(CSRC, "device")
    dev_hdrs = sorted(os.path.join(dev, f) for f in os.listdir(dev))
    key = hashlib.sha1(dsl_source.encode()).hexdigest()[:16]
    pm = os.path.join(BUILD, f"m_{key}.pmdsl")
    cu = os.path.join(BUILD, f"m_{key}.cpp")
    so = os.path.join(BUILD, f"m_{key}.so")
    if not os.path.exists(pm):
        with open(pm, "w") as f:
            f.write(dsl_source)
    if _newer(cu, [pm, DSLC]):
        with open(cu, "w") as f:
            f.write(_run([DSLC, pm]))
    objs = _common_objects()
    if _newer(so, [cu, os.path.join(HERE, "cuda_shim.h")] + dev_hdrs + objs):
        _run([GXX, "-std=c++17", "-O1" if SANITIZE else "-O2", "-fPIC", "-shared", "-w"] + SAN_FLAGS + ["-include", os.path.join(HERE, "cuda_shim.h"), "-I", dev, cu] + objs + ["-ldl", "-o", so + ".tmp"])
        os.replace(so + ".tmp", so)
    return so


_OPS = {"bolus": "B", "infusion": "I", "observation": "O", "missing_observation": "M", "censored_observation": "C",
        "observation_with_error": "E", "covariate": "V", "repeat": "R", "reset": "X", "covariate_fixed": "F"}
_CENS = {None: 0, "none": 0, "None": 0, "bloq": 1, "BLOQ": 1, "aloq": 2, "ALOQ": 2, 0: 0, 1: 1, 2: 2}


def ops_text(subjects):
    """[(id, ops)] -> the line format hostsim_api.cpp parses."""
    out = []
    for sid, ops in subjects:
        out.append(f"S {str(sid).replace(' ', '_') or 'x'}")
        for op in ops:
            k = op[0]
            if k == "bolus":
                out.append(f"B {op[1]!r} {op[2]!r} {op[3]}")
            elif k == "infusion":
                out.append(f"I {op[1]!r} {op[2]!r} {op[3]} {op[4]!r}")
            elif k == "observation":
                out.append(f"O {op[1]!r} {op[2]!r} {op[3]}")
            elif k == "missing_observation":
                out.append(f"M {op[1]!r} {op[2]}")
            elif k == "censored_observation":
                out.append(f"C {op[1]!r} {op[2]!r} {op[3]} {_CENS[op[4]]}")
            elif k == "observation_with_error":
                c = op[4]
                out.append(f"E {op[1]!r} {op[2]!r} {op[3]} {c[0]!r} {c[1]!r} {c[2]!r} {c[3]!r} {_CENS[op[5]]}")
            elif k == "covariate":
                out.append(f"V {op[1]} {op[2]!r} {op[3]!r}")
            elif k == "repeat":
                out.append(f"R {int(op[1])} {float(op[2])!r}")
            elif k == "reset":
                out.append("X")
            elif k == "covariate_fixed":
                out.append(f"F {int(op[1])} {op[2]} {int(bool(op[3]))}")
            else:
                raise ValueError(op)
    return "\n".join(out).replace("np.float64(", "").replace(")", "") + "\n"


SOLVERS = {"Dopri5": 0, "Tsit45": 1, "Sdirk4": 2, "TrBdf2": 3, "Rodas4": 4, "Bdf": 5, "Esdirk34": 6}


class HostSim:
    def __init__(self, dsl_source):
        self.lib = C.CDLL(build_module(dsl_source), mode=C.RTLD_GLOBAL)
        L = self.lib
        L.hs_new.restype = C.c_void_p
        L.hs_new.argtypes = [C.c_char_p]
        L.hs_error.restype = C.c_char_p
        L.hs_error.argtypes = [C.c_void_p]
        L.hs_set_data.argtypes = [C.c_void_p, C.c_char_p]
        L.hs_nobs.restype = C.c_long
        L.hs_nobs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.hs_free.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        L.hs_run.argtypes = [C.c_void_p, C.c_int, dp, C.c_long, C.c_int, dp, C.c_int, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_long),
                             C.POINTER(C.c_ulonglong)]
        self.h = C.c_void_p(L.hs_new(dsl_source.encode()))
        err = L.hs_error(self.h)
        if err:
            raise RuntimeError(err.decode())
        self.nsub = 0

    def set_subjects(self, subjects):
        """subjects: [(id, ops)] as in benches/workloads.py."""
        if self.lib.hs_set_data(self.h, ops_text(subjects).encode()) != 0:
            raise RuntimeError(self.lib.hs_error(self.h).decode())
        self.nsub = len(subjects)
        return self

    def run(self, support_points, error_models=None, solver="Dopri5", rtol=1e-4, atol=1e-4, cov_time=0, max_steps=200000, want_pred=False,
            particles=1, seed=0x5EED, sde_mode=0, em_mode=0, em_dt=0.05, sde_normals=0):
        """error_models: list per output of None | (kind:int 1 additive / 2 proportional, factor, (c0..c3)).
        Returns (psi F-order (nsub, nspp), predictions (nobs, nspp) | None, info)."""
        spp = np.ascontiguousarray(support_points, dtype=np.float64)
        nspp, npar = spp.shape
        ems = np.zeros((max(len(error_models or []), 1), 6))
        for i, m in enumerate(error_models or []):
            if m is not None:
                ems[i] = [m[0], m[1], *m[2]]
        nem = len(error_models or [])
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        out = np.full((self.nsub, nspp), np.nan, order="F") if nem else None
        pred = None
        if want_pred:
            nobs = self.lib.hs_nobs(self.h, None, 0)
            pred = np.full((nobs, nspp), np.nan)
        opts = np.array([rtol, atol, float(cov_time), float(max_steps), 0.0, float(particles), float(seed), float(sde_mode), float(em_mode), float(em_dt),
                         float(sde_normals)])
        code, pair = C.c_int(0), C.c_long(-1)
        counters = (C.c_ulonglong * 4)()
        rc = self.lib.hs_run(self.h, SOLVERS[solver] if isinstance(solver, str) else int(solver), dp(spp), nspp, npar, dp(ems), nem, dp(opts),
                             dp(out) if out is not None else None, dp(pred) if pred is not None else None, C.byref(code), C.byref(pair), counters)
        if rc != 0:
            raise RuntimeError(f"hostsim rc {rc}: {self.lib.hs_error(self.h).decode()}")
        return out, pred, {"code": code.value, "pair": pair.value, "steps": counters[0], "rejected": counters[1], "evals": counters[2], "newton": counters[3]}

    def __del__(self):
        try:
            self.lib.hs_free(self.h)
        except Exception:
            pass
