"""ctypes binding of ``libpharmsol_cuda.so`` — every call goes through the C ABI declared in
``include/pharmsol_cuda.h``.  No torch types cross this boundary (raw pointers and sizes only).

The product never imports ``oracle``; if the shared library is missing the import fails loudly
(there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpharmsol_cuda.so")

ERROR_NAMES = {
    0: "Ok", 1: "NonFiniteLikelihood", 2: "NegativeSigma", 3: "NonFiniteSigma", 4: "InvalidOutputEquation",
    5: "NoneErrorModel", 6: "MissingErrorModel", 7: "SolverFailure", 8: "InputOutOfRange", 9: "OuteqOutOfRange",
    10: "UnknownInputLabel", 11: "UnknownOutputLabel", 12: "ImaginaryRoots", 13: "UnsupportedInputRouteKind",
    14: "MissingCovariate", 15: "OtherError", 64: "CudaError", 65: "CompileError", 66: "InvalidArgument",
}


class PharmsolError(RuntimeError):
    """Mirror of ``pharmsol::PharmsolError`` (src/error/mod.rs:14-49): ``code`` is the C-ABI status,
    ``variant`` the Rust variant name, ``pair`` the failing ``i + j*nsub`` when known."""

    def __init__(self, code, message="", pair=None):
        self.code = int(code)
        self.variant = ERROR_NAMES.get(self.code, f"Error{code}")
        self.pair = pair
        super().__init__(f"{self.variant}: {message}" if message else self.variant)


class pcu_residual_error_model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


class pcu_error_model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("factor", C.c_double),
                ("c0", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("c3", C.c_double)]


def declared_symbols():
    """Every function name declared in include/pharmsol_cuda.h (parsed from the header)."""
    import re
    hdr = os.path.join(_HERE, "..", "include", "pharmsol_cuda.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pharmsol_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pharmsol_b200.build` (nvcc, sm_100a). "
            "pharmsol_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, d, i32, i64, cs, sz = C.c_void_p, C.c_double, C.c_int32, C.c_int64, C.c_char_p, C.c_size_t
    dp = C.POINTER(C.c_double)
    P = C.POINTER
    sig = {
        "pharmsol_cuda_abi_version": (i32, []),
        "pharmsol_cuda_device_count": (i32, [P(i32)]),
        "pharmsol_cuda_ctx_create": (i32, [i32, P(vp)]),
        "pharmsol_cuda_ctx_create_multi": (i32, [P(i32), i32, P(vp)]),
        "pharmsol_cuda_ctx_num_devices": (i32, [vp]),
        "pharmsol_cuda_ctx_device_id": (i32, [vp, i32]),
        "pharmsol_cuda_ctx_num_lanes": (i32, [vp]),
        "pharmsol_cuda_ctx_destroy": (None, [vp]),
        "pharmsol_cuda_last_error_message": (cs, []),
        "pharmsol_cuda_launch_count": (i64, [vp]),
        "pharmsol_cuda_last_kernel_ms": (d, [vp]),
        "pharmsol_cuda_last_counters": (i32, [vp, P(C.c_uint64)]),
        "pharmsol_cuda_host_alloc": (i32, [sz, P(vp)]),
        "pharmsol_cuda_host_free": (i32, [vp]),
        "pharmsol_subject_builder_new": (vp, [cs]),
        "pharmsol_subject_builder_bolus": (None, [vp, d, d, cs]),
        "pharmsol_subject_builder_infusion": (None, [vp, d, d, cs, d]),
        "pharmsol_subject_builder_observation": (None, [vp, d, d, cs]),
        "pharmsol_subject_builder_censored_observation": (None, [vp, d, d, cs, i32]),
        "pharmsol_subject_builder_missing_observation": (None, [vp, d, cs]),
        "pharmsol_subject_builder_observation_with_error": (None, [vp, d, d, cs, d, d, d, d, i32]),
        "pharmsol_subject_builder_covariate": (None, [vp, cs, d, d]),
        "pharmsol_subject_builder_repeat": (None, [vp, i64, d]),
        "pharmsol_subject_builder_reset": (None, [vp]),
        "pharmsol_subject_builder_build": (vp, [vp]),
        "pharmsol_subject_set_covariate_fixed": (i32, [vp, i32, cs, i32]),
        "pharmsol_subject_free": (None, [vp]),
        "pharmsol_data_new": (vp, []),
        "pharmsol_data_add_subject": (i32, [vp, vp]),
        "pharmsol_data_len": (i64, [vp]),
        "pharmsol_data_free": (None, [vp]),
        "pharmsol_data_read_pmetrics": (i32, [cs, P(vp)]),
        "pharmsol_data_from_pmetrics_text": (i32, [cs, sz, P(vp)]),
        "pharmsol_data_describe_json": (i64, [vp, C.c_char_p, sz]),
        "pharmsol_data_expand": (i32, [vp, d, d, P(vp)]),
        "pharmsol_cuda_model_from_dsl": (i32, [vp, cs, sz, P(vp)]),
        "pharmsol_cuda_model_destroy": (None, [vp]),
        "pharmsol_cuda_model_kind": (i32, [vp]),
        "pharmsol_cuda_model_nparams": (i32, [vp]),
        "pharmsol_cuda_model_nstates": (i32, [vp]),
        "pharmsol_cuda_model_nouteqs": (i32, [vp]),
        "pharmsol_cuda_model_info_json": (cs, [vp]),
        "pharmsol_cuda_model_cuda_source": (cs, [vp]),
        "pharmsol_cuda_model_id": (cs, [vp]),
        "pharmsol_cuda_model_set_solver": (i32, [vp, i32, d, d]),
        "pharmsol_cuda_model_set_max_steps": (i32, [vp, i32]),
        "pharmsol_cuda_model_set_particles": (i32, [vp, C.c_uint32, C.c_uint64, i32, i32, d]),
        "pharmsol_cuda_model_set_cov_time": (i32, [vp, i32]),
        "pharmsol_cuda_model_set_sde_normals": (i32, [vp, i32]),
        "pharmsol_cuda_model_compile": (i32, [vp, vp, P(i32)]),
        "pharmsol_cuda_model_precompile_to_cache": (i32, [vp, i32]),
        "pharmsol_cuda_population_create": (i32, [vp, vp, vp, P(pcu_error_model), i32, P(vp)]),
        "pharmsol_cuda_population_set_error_models": (i32, [vp, P(pcu_error_model), i32]),
        "pharmsol_cuda_population_destroy": (None, [vp]),
        "pharmsol_cuda_population_nsubjects": (i64, [vp]),
        "pharmsol_cuda_population_nobservations": (i64, [vp]),
        "pharmsol_cuda_population_obs_offsets": (i32, [vp, P(i64)]),
        "pharmsol_cuda_population_device_bytes": (i64, [vp]),
        "pharmsol_cuda_population_observation_table": (i32, [vp, dp, dp, P(i32), P(i32), P(i32)]),
        "pharmsol_cuda_log_likelihood_matrix": (i32, [vp, vp, vp, dp, i64, i32, dp, P(i32), P(i64)]),
        "pharmsol_cuda_log_likelihood_matrix_device": (i32, [vp, vp, vp, vp, i64, i64, vp, i64, i64, vp]),
        "pharmsol_cuda_log_likelihood_matrix_peers": (i32, [vp, vp, vp, vp, i64, i64, P(vp), i32, i64, i64, vp]),
        "pharmsol_cuda_log_likelihood_matrix_push": (i32, [vp, vp, vp, vp, i64, i64, P(vp), i32, i32, i64, i64, vp]),
        "pharmsol_cuda_log_likelihood_matrix_replicated": (i32, [vp, vp, vp, dp, i64, i32, i32, P(vp), P(i32), P(i64)]),
        "pharmsol_cuda_collect_errors": (i32, [vp, P(i32), P(i64)]),
        "pharmsol_cuda_status_batch_begin": (i32, [vp, vp]),
        "pharmsol_cuda_upload_support_points": (i32, [vp, dp, i64, i32, vp, i64, vp]),
        "pharmsol_cuda_predictions": (i32, [vp, vp, vp, dp, i64, i32, dp]),
        "pharmsol_cuda_predictions_device": (i32, [vp, vp, vp, vp, i64, i64, vp, i64, vp, i64, vp]),
        "pharmsol_cuda_psi": (i32, [vp, vp, vp, dp, i64, i32, dp, P(i32), P(i64)]),
        "pharmsol_cuda_log_likelihood_batch": (i32, [vp, vp, vp, dp, i64, i32, P(pcu_residual_error_model), i32, dp]),
        "pharmsol_cuda_measure_fp64_peak": (i32, [vp, dp, dp]),
        "pharmsol_cuda_model_export_artifact": (i32, [vp, C.c_char_p, P(i32), i32]),
        "pharmsol_cuda_model_host_source": (cs, [vp]),
        "pharmsol_cuda_model_export_host_artifact": (i32, [vp, C.c_char_p]),
        "pharmsol_cuda_model_load_artifact": (i32, [vp, C.c_char_p, P(vp)]),
        "pharmsol_cuda_artifact_info_json": (i64, [C.c_char_p, C.c_char_p, C.c_size_t]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


def _msg():
    m = lib().pharmsol_cuda_last_error_message()
    return m.decode(errors="replace") if m else ""


def check(rc, pair=None):
    if rc != 0:
        raise PharmsolError(rc, _msg(), pair)


def _b(s):
    return str(s).encode()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def artifact_info(path):
    import json
    need = lib().pharmsol_cuda_artifact_info_json(str(path).encode(), None, 0)
    if need < 0:
        raise PharmsolError(15, _msg())
    buf = C.create_string_buffer(need + 1)
    lib().pharmsol_cuda_artifact_info_json(str(path).encode(), buf, need + 1)
    return json.loads(buf.value.decode())


def device_count():
    n = C.c_int32(0)
    check(lib().pharmsol_cuda_device_count(C.byref(n)))
    return n.value


class Context:
    """One device, or — `devices=[...]` — one context driving several devices from this process
    (pharmsol_cuda_ctx_create_multi: column shards, per-device copies into the caller's matrix)."""

    def __init__(self, device=0, devices=None):
        self.ptr = C.c_void_p()
        if devices is not None:
            ids = [int(x) for x in devices]
            arr = (C.c_int32 * len(ids))(*ids)
            check(lib().pharmsol_cuda_ctx_create_multi(arr, len(ids), C.byref(self.ptr)))
            self.devices = ids
            self.device = ids[0]
        else:
            check(lib().pharmsol_cuda_ctx_create(int(device), C.byref(self.ptr)))
            self.device = int(device)
            self.devices = [self.device]

    @property
    def num_devices(self):
        return lib().pharmsol_cuda_ctx_num_devices(self.ptr)

    @property
    def num_lanes(self):
        return lib().pharmsol_cuda_ctx_num_lanes(self.ptr)

    def close(self):
        if self.ptr:
            lib().pharmsol_cuda_ctx_destroy(self.ptr)
            self.ptr = C.c_void_p()

    @property
    def launch_count(self):
        return lib().pharmsol_cuda_launch_count(self.ptr)

    @property
    def last_kernel_ms(self):
        return lib().pharmsol_cuda_last_kernel_ms(self.ptr)

    @property
    def last_counters(self):
        out = (C.c_uint64 * 4)()
        check(lib().pharmsol_cuda_last_counters(self.ptr, out))
        return {"steps": out[0], "rejected": out[1], "evals": out[2], "newton": out[3]}

    def measure_fp64_peak(self):
        t, clk = C.c_double(), C.c_double()
        check(lib().pharmsol_cuda_measure_fp64_peak(self.ptr, C.byref(t), C.byref(clk)))
        return t.value, clk.value

    def collect_errors(self):
        code, pair = C.c_int32(0), C.c_int64(-1)
        rc = lib().pharmsol_cuda_collect_errors(self.ptr, C.byref(code), C.byref(pair))
        if rc != 0:
            raise PharmsolError(rc, _msg(), pair.value)


_contexts = {}


def context(device=0):
    """One shared Context per device — or per device LIST (a multi-device context, e.g. ``context([0, 1, 2, 3])``)."""
    if isinstance(device, (list, tuple)):
        key = tuple(int(x) for x in device)
        if key not in _contexts:
            _contexts[key] = Context(devices=list(key))
        return _contexts[key]
    device = int(device)
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


CENSOR = {None: 0, "none": 0, "None": 0, "bloq": 1, "BLOQ": 1, "aloq": 2, "ALOQ": 2, 0: 0, 1: 1, 2: 2}


class NativeSubject:
    """Builds a pcu_subject from builder ops (see Subject.ops in api.py)."""

    def __init__(self, id, ops):
        L = lib()
        b = L.pharmsol_subject_builder_new(_b(id))
        fixed = []
        for op in ops:
            k = op[0]
            if k == "bolus":
                L.pharmsol_subject_builder_bolus(b, op[1], op[2], _b(op[3]))
            elif k == "infusion":
                L.pharmsol_subject_builder_infusion(b, op[1], op[2], _b(op[3]), op[4])
            elif k == "observation":
                L.pharmsol_subject_builder_observation(b, op[1], op[2], _b(op[3]))
            elif k == "missing_observation":
                L.pharmsol_subject_builder_missing_observation(b, op[1], _b(op[2]))
            elif k == "censored_observation":
                L.pharmsol_subject_builder_censored_observation(b, op[1], op[2], _b(op[3]), CENSOR[op[4]])
            elif k == "observation_with_error":
                c = op[4]
                L.pharmsol_subject_builder_observation_with_error(b, op[1], op[2], _b(op[3]), c[0], c[1], c[2], c[3], CENSOR[op[5]])
            elif k == "covariate":
                L.pharmsol_subject_builder_covariate(b, _b(op[1]), op[2], op[3])
            elif k == "repeat":
                L.pharmsol_subject_builder_repeat(b, int(op[1]), float(op[2]))
            elif k == "reset":
                L.pharmsol_subject_builder_reset(b)
            elif k == "covariate_fixed":
                fixed.append(op)
            else:
                raise ValueError(f"unknown subject op {op!r}")
        self.ptr = C.c_void_p(L.pharmsol_subject_builder_build(b))
        if not self.ptr:
            raise PharmsolError(66, _msg())
        for op in fixed:
            check(L.pharmsol_subject_set_covariate_fixed(self.ptr, int(op[1]), _b(op[2]), int(bool(op[3]))))

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.pharmsol_subject_free(self.ptr)
            self.ptr = None


class NativeData:
    def __init__(self, subjects=None, ptr=None):
        L = lib()
        if ptr is not None:
            self.ptr = ptr
            return
        self.ptr = C.c_void_p(L.pharmsol_data_new())
        for s in subjects:
            ns = NativeSubject(s.id, s.ops)
            check(L.pharmsol_data_add_subject(self.ptr, ns.ptr))

    @classmethod
    def from_pmetrics(cls, path=None, text=None):
        ptr = C.c_void_p()
        if text is not None:
            raw = text.encode() if isinstance(text, str) else bytes(text)
            check(lib().pharmsol_data_from_pmetrics_text(raw, len(raw), C.byref(ptr)))
        else:
            check(lib().pharmsol_data_read_pmetrics(_b(path), C.byref(ptr)))
        return cls(ptr=ptr)

    def expand(self, idelta, tad):
        ptr = C.c_void_p()
        check(lib().pharmsol_data_expand(self.ptr, float(idelta), float(tad), C.byref(ptr)))
        return NativeData(ptr=ptr)

    def describe(self):
        import json
        n = lib().pharmsol_data_describe_json(self.ptr, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().pharmsol_data_describe_json(self.ptr, buf, n + 1)
        return json.loads(buf.value.decode())

    def __len__(self):
        return lib().pharmsol_data_len(self.ptr)

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.pharmsol_data_free(self.ptr)
            self.ptr = None


class Model:
    def __init__(self, ptr):
        self.ptr = ptr

    @classmethod
    def from_dsl(cls, source):
        src = source.encode()
        ptr = C.c_void_p()
        check(lib().pharmsol_cuda_model_from_dsl(None, src, len(src), C.byref(ptr)))
        return cls(ptr)

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.pharmsol_cuda_model_destroy(self.ptr)
            self.ptr = None

    kind = property(lambda self: lib().pharmsol_cuda_model_kind(self.ptr))
    nparams = property(lambda self: lib().pharmsol_cuda_model_nparams(self.ptr))
    nstates = property(lambda self: lib().pharmsol_cuda_model_nstates(self.ptr))
    nouteqs = property(lambda self: lib().pharmsol_cuda_model_nouteqs(self.ptr))
    id = property(lambda self: lib().pharmsol_cuda_model_id(self.ptr).decode())
    cuda_source = property(lambda self: lib().pharmsol_cuda_model_cuda_source(self.ptr).decode())

    @property
    def info(self):
        import json
        return json.loads(lib().pharmsol_cuda_model_info_json(self.ptr).decode())

    def set_solver(self, solver, rtol, atol):
        check(lib().pharmsol_cuda_model_set_solver(self.ptr, int(solver), float(rtol), float(atol)))

    def set_max_steps(self, n):
        check(lib().pharmsol_cuda_model_set_max_steps(self.ptr, int(n)))

    def set_particles(self, n, seed=0, sde_mode=0, em_mode=0, em_dt=0.0):
        check(lib().pharmsol_cuda_model_set_particles(self.ptr, int(n), int(seed), int(sde_mode), int(em_mode), float(em_dt)))

    def set_sde_normals(self, precision):
        check(lib().pharmsol_cuda_model_set_sde_normals(self.ptr, int(precision)))

    def set_cov_time(self, mode):
        check(lib().pharmsol_cuda_model_set_cov_time(self.ptr, int(mode)))

    def compile(self, ctx):
        src = C.c_int32(-1)
        check(lib().pharmsol_cuda_model_compile(ctx.ptr, self.ptr, C.byref(src)))
        return {0: "aot", 1: "cubin-cache", 2: "nvrtc", 3: "artifact"}[src.value]

    def export_artifact(self, path, solvers=()):
        arr = (C.c_int32 * max(len(solvers), 1))(*[int(s) for s in solvers])
        check(lib().pharmsol_cuda_model_export_artifact(self.ptr, str(path).encode(), arr, len(solvers)))
        return str(path)

    @property
    def host_source(self):
        """C++ source of the host twin exporting the reference's frozen compiled-backend symbols."""
        return lib().pharmsol_cuda_model_host_source(self.ptr).decode()

    def export_host_artifact(self, path):
        check(lib().pharmsol_cuda_model_export_host_artifact(self.ptr, str(path).encode()))
        return str(path)

    @classmethod
    def from_artifact(cls, path):
        ptr = C.c_void_p()
        check(lib().pharmsol_cuda_model_load_artifact(None, str(path).encode(), C.byref(ptr)))
        return cls(ptr)

    def precompile_to_cache(self, solver=0):
        check(lib().pharmsol_cuda_model_precompile_to_cache(self.ptr, int(solver)))


def _em_array(error_models):
    """error_models: list of None | (kind:int, factor, (c0,c1,c2,c3)) per output equation."""
    n = len(error_models)
    arr = (pcu_error_model * max(n, 1))()
    for i, m in enumerate(error_models):
        if m is None:
            arr[i].kind = 0
            continue
        kind, factor, poly = m
        arr[i].kind = int(kind)
        arr[i].factor = float(factor)
        arr[i].c0, arr[i].c1, arr[i].c2, arr[i].c3 = [float(c) for c in poly]
    return arr, n


class Population:
    def __init__(self, ctx, model, data, error_models=None):
        self.ptr = C.c_void_p()
        self.ctx = ctx
        if error_models:
            arr, n = _em_array(error_models)
        else:
            arr, n = None, 0
        check(lib().pharmsol_cuda_population_create(ctx.ptr, model.ptr, data.ptr, arr, n, C.byref(self.ptr)))

    def set_error_models(self, error_models):
        if error_models:
            arr, n = _em_array(error_models)
        else:
            arr, n = None, 0
        check(lib().pharmsol_cuda_population_set_error_models(self.ptr, arr, n))

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.pharmsol_cuda_population_destroy(self.ptr)
            self.ptr = None

    nsubjects = property(lambda self: lib().pharmsol_cuda_population_nsubjects(self.ptr))
    nobservations = property(lambda self: lib().pharmsol_cuda_population_nobservations(self.ptr))
    device_bytes = property(lambda self: lib().pharmsol_cuda_population_device_bytes(self.ptr))

    def observation_table(self):
        """time, value (NaN = missing), outeq, occasion, censoring per prediction row."""
        n = self.nobservations
        t, v = np.empty(n), np.empty(n)
        oq, oc, ce = (np.empty(n, dtype=np.int32) for _ in range(3))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        check(lib().pharmsol_cuda_population_observation_table(self.ptr, _dp(t), _dp(v), ip(oq), ip(oc), ip(ce)))
        return t, v, oq, oc, ce

    def obs_offsets(self):
        out = (C.c_int64 * (self.nsubjects + 1))()
        check(lib().pharmsol_cuda_population_obs_offsets(self.ptr, out))
        return np.array(out[:], dtype=np.int64)


def log_likelihood_matrix(ctx, model, pop, support_points, out=None, exponentiate=False):
    """Host-buffer call: row-major (nspp, P) in, F-order (nsub, nspp) out."""
    spp = np.ascontiguousarray(support_points, dtype=np.float64)
    if spp.ndim != 2:
        raise PharmsolError(66, "support_points must be 2-D (rows = support points)")
    nspp, npar = spp.shape
    nsub = pop.nsubjects
    if out is None:
        out = np.empty((nsub, nspp), dtype=np.float64, order="F")
    code, pair = C.c_int32(0), C.c_int64(-1)
    fn = lib().pharmsol_cuda_psi if exponentiate else lib().pharmsol_cuda_log_likelihood_matrix
    rc = fn(ctx.ptr, model.ptr, pop.ptr, _dp(spp), nspp, npar, _dp(out), C.byref(code), C.byref(pair))
    if rc != 0:
        raise PharmsolError(rc, _msg(), pair.value if pair.value >= 0 else None)
    return out


def predictions(ctx, model, pop, support_points):
    """(nobs_total, nspp) row-major predictions for every pair."""
    spp = np.ascontiguousarray(support_points, dtype=np.float64)
    nspp, npar = spp.shape
    out = np.empty((pop.nobservations, nspp), dtype=np.float64)
    check(lib().pharmsol_cuda_predictions(ctx.ptr, model.ptr, pop.ptr, _dp(spp), nspp, npar, _dp(out)))
    return out


# ---- device-resident entry points (raw device pointers; the caller owns the memory) ---------------
def upload_support_points(ctx, support_points, spp_soa_ptr, ld_spp, stream=0):
    """Row-major host (nspp, P) -> parameter-major SoA device buffer spp[k*ld_spp + j]."""
    spp = np.ascontiguousarray(support_points, dtype=np.float64)
    nspp, npar = spp.shape
    check(lib().pharmsol_cuda_upload_support_points(ctx.ptr, _dp(spp), nspp, npar, C.c_void_p(int(spp_soa_ptr)), int(ld_spp),
                                                    C.c_void_p(int(stream) or None)))


def status_batch_begin(ctx, stream=0):
    """The launches that follow share one error word / counter set until the next collect_errors."""
    check(lib().pharmsol_cuda_status_batch_begin(ctx.ptr, C.c_void_p(int(stream) or None)))


def log_likelihood_matrix_device(ctx, model, pop, spp_soa_ptr, ncols, ld_spp, out_ptr, ld_out, first_col=0, stream=0):
    """Asynchronous launch with everything resident in HBM: out[i + j*ld_out] for j < ncols."""
    check(lib().pharmsol_cuda_log_likelihood_matrix_device(ctx.ptr, model.ptr, pop.ptr, C.c_void_p(int(spp_soa_ptr)), int(ncols),
                                                           int(ld_spp), C.c_void_p(int(out_ptr)), int(ld_out), int(first_col),
                                                           C.c_void_p(int(stream) or None)))


def host_alloc(nbytes):
    p = C.c_void_p()
    check(lib().pharmsol_cuda_host_alloc(int(nbytes), C.byref(p)))
    return p.value


def host_free(ptr):
    check(lib().pharmsol_cuda_host_free(C.c_void_p(int(ptr))))


def pinned_array(shape, order="C"):
    """numpy float64 array backed by page-locked host memory (cudaMallocHost through the C ABI)."""
    n = int(np.prod(shape))
    ptr = host_alloc(max(n, 1) * 8)
    buf = (C.c_double * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape, order=order)
    return arr, ptr


def log_likelihood_batch(ctx, model, pop, parameters, residual_models):
    """likelihood/mod.rs:119-177: one parameter row per subject -> nsub log-likelihoods.
    residual_models: list per outeq of None | (kind:int, a, b)."""
    prm = np.ascontiguousarray(parameters, dtype=np.float64)
    if prm.ndim != 2:
        raise PharmsolError(66, "parameters must be 2-D (rows = subjects)")
    n = len(residual_models)
    arr = (pcu_residual_error_model * max(n, 1))()
    for k, m in enumerate(residual_models):
        if m is not None:
            arr[k].kind, arr[k].a, arr[k].b = int(m[0]), float(m[1]), float(m[2])
    out = np.empty(pop.nsubjects, dtype=np.float64)
    check(lib().pharmsol_cuda_log_likelihood_batch(ctx.ptr, model.ptr, pop.ptr, _dp(prm), prm.shape[0], prm.shape[1], arr, n, _dp(out)))
    return out


def log_likelihood_matrix_peers(ctx, model, pop, spp_soa_ptr, ncols, ld_spp, peer_ptrs, ld_out, first_col, stream=0):
    """Asynchronous launch that stores every result into the full psi of every rank (fused all-gather)."""
    arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(q)) for q in peer_ptrs])
    check(lib().pharmsol_cuda_log_likelihood_matrix_peers(ctx.ptr, model.ptr, pop.ptr, C.c_void_p(int(spp_soa_ptr)), int(ncols), int(ld_spp),
                                                          arr, len(peer_ptrs), int(ld_out), int(first_col), C.c_void_p(int(stream) or None)))


def log_likelihood_matrix_push(ctx, model, pop, spp_soa_ptr, ncols, ld_spp, peer_ptrs, self_index, ld_out, first_col, stream=0):
    """Asynchronous chunked launch into peer_ptrs[self_index] (a full matrix) + copy-engine pushes of every finished
    chunk to the other ranks' matrices (pharmsol_cuda_log_likelihood_matrix_push)."""
    arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(q)) for q in peer_ptrs])
    check(lib().pharmsol_cuda_log_likelihood_matrix_push(ctx.ptr, model.ptr, pop.ptr, C.c_void_p(int(spp_soa_ptr)), int(ncols), int(ld_spp),
                                                         arr, len(peer_ptrs), int(self_index), int(ld_out), int(first_col),
                                                         C.c_void_p(int(stream) or None)))


GATHER_COPY_ENGINE, GATHER_PEER_STORES = 0, 1


def log_likelihood_matrix_replicated(ctx, model, pop, support_points, gather=GATHER_COPY_ENGINE):
    """Multi-device context: host support points in, the whole psi resident on every device out.
    Returns the list of device pointers (one column-major (nsub, nspp) matrix per device, library-owned)."""
    spp = np.ascontiguousarray(support_points, dtype=np.float64)
    nspp, npar = spp.shape
    n = ctx.num_devices
    ptrs = (C.c_void_p * n)()
    code, pair = C.c_int32(0), C.c_int64(-1)
    rc = lib().pharmsol_cuda_log_likelihood_matrix_replicated(ctx.ptr, model.ptr, pop.ptr, _dp(spp), nspp, npar, int(gather), ptrs, C.byref(code), C.byref(pair))
    if rc != 0:
        raise PharmsolError(rc, _msg(), pair.value if pair.value >= 0 else None)
    return [int(q or 0) for q in ptrs]
