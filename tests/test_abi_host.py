"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol
``include/pharmsol_cuda.h`` declares, the DSL front end / CUDA-C emitter behave like the reference's
pharmsol-dsl on the reference's own sources, data building mirrors SubjectBuilder, and every compute
entry point FAILS LOUDLY without a device (there is no CPU fallback).  No compute calls here."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import fixtures as FX
from conftest import ROOT, has_gpu


def test_library_exports_every_declared_symbol(libpath):
    from pharmsol_b200 import _lib
    L = C.CDLL(libpath)
    declared = _lib.declared_symbols()
    assert len(declared) >= 50
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    # and the binding table covers the header exactly
    assert sorted(_lib.lib()._signatures) == declared
    assert L.pharmsol_cuda_abi_version() == 2


def test_no_torch_types_in_header():
    hdr = open(os.path.join(ROOT, "include", "pharmsol_cuda.h")).read()
    assert "torch" not in hdr.lower() and "at::" not in hdr and "std::" not in hdr
    assert 'extern "C"' in hdr


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "pharmsol_cuda.h"\nint main(void){return pharmsol_cuda_abi_version()==PHARMSOL_CUDA_ABI_VERSION?0:1;}\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_compute_fails_loudly_without_device(ps):
    """No CPU fallback: creating a context (a prerequisite of every compute call) must fail with
    PCU_ERR_CUDA, never silently succeed."""
    from pharmsol_b200 import _lib
    with pytest.raises(ps.PharmsolError) as e:
        _lib.Context(0)
    assert e.value.code == 64
    eq = ps.Equation.from_dsl(FX.ANALYTICAL_SOURCE)
    data = ps.Data([ps.Subject("s", FX.CORPUS["analytical"][3])])
    with pytest.raises(ps.PharmsolError) as e:
        eq.estimate_predictions(data.subjects[0], FX.CORPUS["analytical"][2])
    assert e.value.code == 64


def test_bad_arguments_become_errors_not_crashes(ps):
    """Nothing may unwind (or overflow the stack) through the C boundary: NULL labels, absurd repeat counts, NULL
    handles and degenerate DSL sources come back as status codes / NULL + a message."""
    from pharmsol_b200 import _lib
    L = _lib.lib()
    for poison in (lambda b: L.pharmsol_subject_builder_bolus(b, 0.0, 1.0, None),
                   lambda b: L.pharmsol_subject_builder_observation(b, 0.0, 1.0, None),
                   lambda b: L.pharmsol_subject_builder_covariate(b, None, 0.0, 1.0),
                   lambda b: (L.pharmsol_subject_builder_bolus(b, 0.0, 1.0, b"iv"), L.pharmsol_subject_builder_repeat(b, -1, 1.0)),
                   lambda b: (L.pharmsol_subject_builder_bolus(b, 0.0, 1.0, b"iv"), L.pharmsol_subject_builder_repeat(b, 10**12, 1.0)),
                   lambda b: L.pharmsol_subject_builder_censored_observation(b, 0.0, 1.0, b"cp", 9)):
        b = L.pharmsol_subject_builder_new(b"bad")
        poison(b)
        L.pharmsol_subject_builder_bolus(b, 1.0, 1.0, b"iv")        # later calls on a poisoned builder are ignored
        assert L.pharmsol_subject_builder_build(b) is None
        assert b"subject `bad`" in L.pharmsol_cuda_last_error_message()
    L.pharmsol_subject_builder_bolus(None, 0.0, 1.0, b"iv")
    assert L.pharmsol_subject_builder_build(None) is None
    with pytest.raises(ps.PharmsolError, match="repeat count -3"):
        _lib.NativeSubject("s", [("bolus", 0.0, 1.0, "iv"), ("repeat", -3, 1.0)])
    assert L.pharmsol_cuda_model_kind(None) == -1 and L.pharmsol_cuda_model_nparams(None) == -1
    assert L.pharmsol_cuda_model_info_json(None) == b"" and L.pharmsol_cuda_model_id(None) == b""
    assert L.pharmsol_data_add_subject(None, None) == 66 and L.pharmsol_data_len(None) == 0
    # degenerate sources: deep nesting and endless operator chains are compile errors, not stack overflows
    base = "name = d\nkind = ode\nparams = ke\nstates = c\noutputs = y\ndx(c) = -ke * c\nout(y) = %s ~ continuous()\n"
    for expr in ("(" * 50000 + "c" + ")" * 50000, "-" * 50000 + "c", "c" + " + c" * 50000, "c" + " ^ c" * 50000,
                 "if (c > 0) { " * 5000 + "1" + " } else { 2 }" * 5000):
        with pytest.raises(ps.PharmsolError, match="too deep|too many chained"):
            ps.Equation.from_dsl(base % expr)
    assert ps.Equation.from_dsl(base % ("c" + " + c" * 1000)).nstates() == 1


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under pharmsol_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pharmsol_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        if any(seg in dirpath for seg in ("_build", "_cubin_cache", "__pycache__")):
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(import|from)\s+oracle\b", text, re.M) or "liboracle" in text or "pharmsol_oracle" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


# ---- DSL front end -------------------------------------------------------------------------------
@pytest.mark.parametrize("case", list(FX.CORPUS))
def test_reference_corpus_sources_compile(ps, case):
    """tests/support/runtime_corpus.rs sources parse, analyse and lower to CUDA C."""
    src, _, p, _, _ = FX.CORPUS[case]
    eq = ps.Equation.from_dsl(src)
    assert eq.nparams() == len(p)
    info = eq.info
    assert info["outputs"] == [{"name": "cp", "index": 0}] and eq.output_names() == ["cp"]
    cu = eq.cuda_source
    assert "PSI_DEFINE_ENTRY" in cu and "psi_engine.cuh" in cu
    assert eq.kind() == {"ode": 0, "ode_full": 0, "analytical": 1, "analytical_full": 1, "sde": 2}[case]


def test_model_info_mirrors_native_model_info(ps):
    eq = ps.Equation.from_dsl(FX.ODE_FULL_SOURCE)
    info = eq.info
    assert info["parameters"] == ["ka", "ke", "kcp", "kpc", "v", "tlag", "f_oral", "base_depot", "base_central", "base_peripheral"]
    # the serde shape of NativeModelInfo (dsl/model_info.rs:17-101), so a Rust caller can deserialize it as is
    assert info["kind"] == "Ode" and info["analytical"] is None and info["particles"] is None
    assert info["covariates"] == [{"name": "wt", "index": 0, "interpolation": "Linear"}, {"name": "renal", "index": 1, "interpolation": "Linear"}]
    assert info["states"] == [{"name": "depot", "offset": 0}, {"name": "central", "offset": 1}, {"name": "peripheral", "offset": 2}]
    assert info["derived"] == ["adjusted_ke", "adjusted_kcp", "adjusted_v"] and info["derived_len"] == 3
    assert (info["state_len"], info["output_len"], info["route_len"]) == (3, 1, 2)
    routes = {r["name"]: r for r in info["routes"]}
    # bolus and infusion routes are numbered independently from 0 (metadata.rs:926-957)
    assert routes["oral"]["index"] == 0 and routes["load"]["index"] == 1 and routes["iv"]["index"] == 0
    assert routes["oral"]["kind"] == "Bolus" and routes["iv"]["kind"] == "Infusion"
    assert [r["declaration_index"] for r in info["routes"]] == [0, 1, 2]
    assert routes["oral"]["destination_name"] == "depot" and routes["iv"]["destination_offset"] == 1
    assert all(r["inject_input_to_destination"] for r in info["routes"])
    ana = ps.Equation.from_dsl(FX.ANALYTICAL_FULL_SOURCE).info
    assert ana["kind"] == "Analytical" and ana["analytical"] == "OneCompartmentWithAbsorption"
    sde = ps.Equation.from_dsl(FX.SDE_SOURCE).info
    assert sde["kind"] == "Sde" and sde["particles"] == 16
    explicit = ps.Equation.from_dsl("name = x\nkind = ode\nparams = ke\nstates = c\noutputs = y\ninfusion(iv) -> c\n"
                                    "dx(c) = -ke * c + 0.5 * rate(iv)\nout(y) = c ~ continuous()\n").info
    assert explicit["routes"][0]["inject_input_to_destination"] is False
    assert routes["oral"]["has_lag"] and routes["oral"]["has_bioavailability"] and not routes["load"]["has_lag"]


def test_canonical_block_form_equals_authoring_form(ps):
    canonical = """
model one_cpt {
  kind ode
  parameters { ke, v }
  states { central }
  routes { iv -> central }
  dynamics { ddt(central) = -ke * central }
  outputs { cp = central / v }
}
"""
    eq = ps.Equation.from_dsl(canonical)
    assert eq.kind() == 0 and eq.nstates() == 1 and eq.output_names() == ["cp"]


def test_precedence_and_typing_rules_in_emitted_code(ps):
    """pharmsol-dsl parser.rs:1242-1261 (unary binds tighter than ^; ^ right-assoc) and analyze.rs
    (integer-valued literals are Int; / and ^ always real)."""
    src = """
name = prec
kind = ode
params = a, b
states = x
outputs = y
bolus(d) -> x
dx(x) = -a^2 - 2^3^2 * x
out(y) = 7 / 2 * x + b ~ continuous()
"""
    cu = ps.Equation.from_dsl(src).cuda_source
    body = cu[cu.index("dynamics("):]
    # -a^2 == (-a)^2
    assert "psi::pow_2((-p[0]))" in body or re.search(r"pow\(\s*\(-p\[0\]\)\s*,", body), body[:400]
    # 2^3^2 == 2^(3^2) (right associative)
    assert "512.0 * x[0]" in body   # constant sub-expressions are pre-folded (ExecutionExpr.constant)
    # 7 / 2 is real division -> 3.5, never integer 3
    out = cu[cu.index("outputs("):]
    assert "3.5 * x[0]" in out, out[:300]


def test_dsl_errors_are_compile_errors(ps):
    with pytest.raises(ps.PharmsolError) as e:
        ps.Equation.from_dsl("name = bad\nkind = ode\nparams = a\nstates = x\noutputs = y\ndx(x) = -a * nope\nout(y) = x ~ continuous()\n")
    assert e.value.code == 65 and "nope" in str(e.value)
    with pytest.raises(ps.PharmsolError) as e:
        ps.Equation.from_dsl("name = bad\nkind = analytical\nparams = ke\nstates = c\noutputs = y\nstructure = no_such_kernel\nout(y) = c ~ continuous()\n")
    assert e.value.code == 65
    # every state must be assigned in dynamics (analyze.rs:2414-2432)
    with pytest.raises(ps.PharmsolError) as e:
        ps.Equation.from_dsl("name = bad\nkind = ode\nparams = a\nstates = x, z\noutputs = y\nbolus(d) -> x\ndx(x) = -a * x\nout(y) = x ~ continuous()\n")
    assert e.value.code == 65


def test_all_twelve_structures_are_known(ps):
    for k in FX.KERNEL_PARAMS:
        eq = ps.Equation.from_dsl(FX.kernel_dsl(k))
        assert eq.kind() == 1 and eq.nparams() == len(FX.KERNEL_PARAMS[k])


def test_symbolic_jacobian_is_emitted_for_ode_models(ps):
    from benches import workloads as W
    cu = ps.Equation.from_dsl(W.model_source("c4_mm_effect")).cuda_source
    assert "jacobian(" in cu and "J[0 * NSTATE + 0]" in cu


def test_nvrtc_compiles_generated_source_without_a_device(ps, tmp_path, monkeypatch):
    """The DSL -> CUDA C -> NVRTC (sm_100a cubin) leg needs no GPU: compile a model that has no
    ahead-of-time twin and check a cubin lands in the cache."""
    monkeypatch.setenv("PHARMSOL_B200_CUBIN_CACHE", str(tmp_path))
    from pharmsol_b200 import _lib
    m = _lib.Model.from_dsl(FX.ODE_SOURCE)
    m.precompile_to_cache(0)
    files = os.listdir(tmp_path)
    assert files and files[0].endswith(".cubin") and os.path.getsize(tmp_path / files[0]) > 10000
    dump = subprocess.run(["cuobjdump", "-elf", str(tmp_path / files[0])], capture_output=True, text=True).stdout
    assert "sm_100" in dump or "EF_CUDA_SM100" in dump or dump == ""


# ---- host data model -----------------------------------------------------------------------------------
def test_subject_builder_mirror(ps):
    s = (ps.Subject.builder("id1").bolus(0.0, 100.0, "oral").infusion(1.0, 50.0, "iv", 0.5)
         .observation(2.0, 1.5, "cp").missing_observation(3.0, "cp").covariate("wt", 0.0, 70.0)
         .repeat(2, 12.0).reset().bolus(0.0, 10.0, "oral").build())
    d = ps.Data([s])
    assert len(d) == 1 and len(d.native()) == 1


def test_parameter_order(ps):
    eq = ps.Equation.from_dsl(FX.ANALYTICAL_SOURCE)
    order = ps.ParameterOrder.with_model(eq, ["v", "ke", "ka", "f_oral", "tlag"])
    spp = np.array([[25.0, 0.15, 1.0, 0.8, 0.5]])
    assert order.matrix(spp).tolist() == [[1.0, 0.15, 25.0, 0.5, 0.8]]
    with pytest.raises(ps.PharmsolError):
        ps.ParameterOrder.with_model(eq, ["v", "ke"])


def test_error_models_bind_by_label(ps):
    eq = ps.Equation.from_dsl(FX.ANALYTICAL_SOURCE)
    ems = ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0))
    assert ems.bound(eq.output_names()) == [(1, 0.0, (0.1, 0.1, 0.0, 0.0))]
    with pytest.raises(ps.PharmsolError) as e:
        ps.AssayErrorModels().add("nope", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0)).bound(eq.output_names())
    assert e.value.code == 11


def test_macro_style_builders_lower_to_the_same_model_as_dsl_text(ps):
    """analytical!{..} / ode!{..} / sde!{..} declarations (pharmsol-macros/src/expand/*.rs) mirrored as builders that take
    DSL expression strings: they must lower to the same compiled model (same id = same generated CUDA) as the text form."""
    from benches import workloads as W
    a = ps.analytical(name="one_cpt_iv", params=["ke", "v"], states=["central"], outputs=["cp"], routes=["infusion(iv) -> central"],
                      structure="one_compartment", out={"cp": "central / v"})
    assert a._model.id == ps.Equation.from_dsl(W.model_source("c1_one_cpt_iv"))._model.id and a.kind() == ps.EqnKind.Analytical
    o = ps.ode(name="two_cpt_oral", params=["ka", "ke", "kcp", "kpc", "v"], states=["depot", "central", "peripheral"], outputs=["cp"],
               routes=["bolus(oral) -> depot"],
               diffeq={"depot": "-ka * depot", "central": "ka * depot - (ke + kcp) * central + kpc * peripheral",
                       "peripheral": "kcp * central - kpc * peripheral"}, out={"cp": "central / v"})
    assert o._model.id == ps.Equation.from_dsl(W.model_source("c2_two_cpt_oral_ode"))._model.id and isinstance(o, ps.ODE)
    s = ps.sde(name="one_cpt_sde", params=["ke", "sigma", "v"], states=["central"], outputs=["cp"],
               routes=["infusion(iv) -> central", "bolus(load) -> central"], drift={"central": "-ke * central"},
               diffusion={"central": "sigma"}, out={"cp": "central / v"}, particles=1000)
    assert s._model.id == ps.Equation.from_dsl(W.model_source("c5_one_cpt_sde"))._model.id and isinstance(s, ps.SDE)


def test_rust_ffi_binds_every_declared_symbol():
    """integration/rust/.../ffi.rs is generated from the header (scripts/gen_rust_ffi.py): it must be up to date and must
    declare every function of include/pharmsol_cuda.h, with pointer-to-pointer arguments spelled out."""
    import re
    import subprocess
    import sys
    from conftest import ROOT
    from pharmsol_b200 import _lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_ffi.py"), "--check"])
    assert r.returncode == 0, "integration/rust/src/simulator/cuda/ffi.rs is stale: run scripts/gen_rust_ffi.py"
    text = open(os.path.join(ROOT, "integration", "rust", "src", "simulator", "cuda", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (pharmsol_[a-z0-9_]+)\(", text))
    assert bound == set(_lib.declared_symbols())
    assert "out_full_peers: *const *mut f64" in text and "dev_out: *mut *mut f64" in text and "device_ids: *const i32" in text
    # the glue uses only functions the FFI block declares
    glue = open(os.path.join(ROOT, "integration", "rust", "src", "simulator", "cuda", "mod.rs")).read()
    used = set(re.findall(r"ffi::(pharmsol_[a-z0-9_]+)", glue))
    assert used and used <= bound, used - bound
