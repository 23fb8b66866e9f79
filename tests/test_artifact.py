"""CUDA-target model artifact (.pkm), SURVEY §8 f.2 — the counterpart of compile_module_source_to_aot /
load_aot_model / read_aot_model_info (reference src/dsl/aot.rs:146-353, tests/dsl_aot_roundtrip-style checks:
export, inspect, reload, reject bad files).  Exporting needs NVRTC only; launching needs a GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

import fixtures as FX

ODE_SRC = ("name = artifact_ode\nkind = ode\nparams = ka, ke, v\nstates = gut, central\noutputs = plasma\nbolus(po) -> gut\n"
           "dx(gut) = -ka * gut\ndx(central) = ka * gut - ke * central * 1.0000001\nout(plasma) = central / v ~ continuous()\n")
ANA_SRC = ("name = artifact_ana\nkind = analytical\nparams = ka, ke, v\nstates = gut, central\noutputs = plasma\nbolus(po) -> gut\n"
           "structure = one_compartment_with_absorption\nout(plasma) = central / v * 1.0000001 ~ continuous()\n")
SDE_SRC = ("name = artifact_sde\nkind = sde\nparams = ke, v, s\nstates = central\noutputs = cp\nparticles = 64\nbolus(iv) -> central\n"
           "dx(central) = -ke * central\nnoise(central) = s\nout(cp) = central / v ~ continuous()\n")


def _sections(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"PKMCUDA\0"
    api, nsec = struct.unpack_from("<II", raw, 8)
    off, out = 16, []
    for _ in range(nsec):
        tag, aux, n = struct.unpack_from("<IIQ", raw, off)
        out.append((tag, aux, raw[off + 16:off + 16 + n]))
        off += 16 + n
    assert off + 8 == len(raw)
    return api, out


def test_export_inspect_reload_without_a_device(ps, tmp_path):
    path = ps.compile_module_source_to_aot(ODE_SRC, tmp_path / "m.pkm", solvers=(ps.OdeSolver.Dopri5, ps.OdeSolver.Rodas4),
                                           configure=lambda e: e.with_solver(ps.OdeSolver.Rodas4).with_tolerances(1e-7, 1e-9))
    api, secs = _sections(path)
    assert api == 1 and [t for t, _, _ in secs] == [1, 2, 3, 4, 5, 5]
    assert secs[1][2].decode() == ODE_SRC
    info = ps.read_aot_model_info(path)
    assert info["format"] == "pharmsol-cuda-pkm" and info["engine_matches"] is True and info["engine"].startswith("sm_100a ")
    assert info["model"]["name"] == "artifact_ode" and info["model"]["parameters"] == ["ka", "ke", "v"]
    assert info["settings"]["solver"] == 4 and info["settings"]["rtol"] == 1e-7 and info["settings"]["atol"] == 1e-9
    assert [k["solver"] for k in info["kernels"]] == [0, 4] and all(k["cubin_bytes"] > 10000 for k in info["kernels"])
    cubin = tmp_path / "k.cubin"
    cubin.write_bytes(secs[5][2])
    dump = subprocess.run(["cuobjdump", "-elf", str(cubin)], capture_output=True, text=True).stdout
    assert dump == "" or "sm_100" in dump or "SM100" in dump
    sym = subprocess.run(["cuobjdump", "-symbols", str(cubin)], capture_output=True, text=True).stdout
    assert sym == "" or info["kernels"][1]["entry"] in sym
    eq = ps.load_aot_model(path)
    assert isinstance(eq, ps.ODE) and (eq._solver, eq._rtol, eq._atol) == (4, 1e-7, 1e-9)
    assert eq.parameter_names() == ["ka", "ke", "v"] and eq.output_names() == ["plasma"]
    assert ps.load_runtime_artifact(path, ps.RuntimeArtifactFormat.CudaAot).nstates() == 2


def test_settings_of_every_model_kind_survive_the_round_trip(ps, tmp_path):
    a = ps.Equation.from_dsl(ANA_SRC).with_cov_time(ps.CovTime.IntervalLength)
    a.export_artifact(tmp_path / "a.pkm")
    assert ps.read_aot_model_info(tmp_path / "a.pkm")["settings"]["cov_time"] == ps.CovTime.IntervalLength
    assert isinstance(ps.load_aot_model(tmp_path / "a.pkm"), ps.Analytical)
    s = ps.Equation.from_dsl(SDE_SRC).with_particles(96).with_seed(1234).with_mode(ps.SdeMode.ParticleFilter).with_stepper(ps.EmMode.FixedStep, 0.02)
    s.export_artifact(tmp_path / "s.pkm")
    back = ps.load_aot_model(tmp_path / "s.pkm")
    assert isinstance(back, ps.SDE) and (back._np, back._seed, back._mode, back._em, back._dt) == (96, 1234, ps.SdeMode.ParticleFilter, ps.EmMode.FixedStep, 0.02)
    assert [k["solver"] for k in ps.read_aot_model_info(tmp_path / "s.pkm")["kernels"]] == [0]


def test_bad_artifacts_are_rejected(ps, tmp_path):
    path = ps.Equation.from_dsl(ANA_SRC).export_artifact(tmp_path / "a.pkm")
    raw = bytearray(open(path, "rb").read())
    flipped = bytearray(raw); flipped[len(raw) // 2] ^= 0x40
    (tmp_path / "flip.pkm").write_bytes(flipped)
    with pytest.raises(ps.PharmsolError, match="checksum"):
        ps.load_aot_model(tmp_path / "flip.pkm")
    (tmp_path / "short.pkm").write_bytes(raw[: len(raw) // 3])
    with pytest.raises(ps.PharmsolError, match="checksum|truncated"):
        ps.load_aot_model(tmp_path / "short.pkm")
    (tmp_path / "other.pkm").write_bytes(b"\x7fELF" + bytes(64))
    with pytest.raises(ps.PharmsolError, match="not a pharmsol CUDA artifact"):
        ps.read_aot_model_info(tmp_path / "other.pkm")
    with pytest.raises(ps.PharmsolError, match="cannot open"):
        ps.load_aot_model(tmp_path / "absent.pkm")
    # another API version (checksum recomputed so only the version differs): aot.rs:395-407 ApiVersionMismatch
    v2 = bytearray(raw[:-8]); struct.pack_into("<I", v2, 8, 2)
    h = 1469598103934665603
    for b in v2:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    (tmp_path / "v2.pkm").write_bytes(bytes(v2) + struct.pack("<Q", h))
    with pytest.raises(ps.PharmsolError, match="API version mismatch: expected 1, found 2"):
        ps.load_aot_model(tmp_path / "v2.pkm")


def test_unknown_sections_are_skipped(ps, tmp_path):
    """Forward compatibility inside one API version: a reader ignores section tags it does not know."""
    path = ps.Equation.from_dsl(ANA_SRC).export_artifact(tmp_path / "a.pkm")
    raw = open(path, "rb").read()
    body = bytearray(raw[:-8])
    struct.pack_into("<I", body, 12, struct.unpack_from("<I", body, 12)[0] + 1)
    extra = b"future section payload"
    body += struct.pack("<IIQ", 99, 7, len(extra)) + extra
    h = 1469598103934665603
    for b in body:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    (tmp_path / "future.pkm").write_bytes(bytes(body) + struct.pack("<Q", h))
    info = ps.read_aot_model_info(tmp_path / "future.pkm")
    assert info["model"]["name"] == "artifact_ana" and len(info["kernels"]) == 1
    assert ps.load_aot_model(tmp_path / "future.pkm").parameter_names() == ["ka", "ke", "v"]


def test_foreign_engine_cubin_is_not_trusted(ps, tmp_path, monkeypatch):
    """An artifact built by another engine build keeps working through its DSL source; its device code is dropped."""
    path = ps.Equation.from_dsl(ANA_SRC).export_artifact(tmp_path / "a.pkm")
    monkeypatch.setenv("PHARMSOL_B200_NVRTC_FLAGS", "-DPSI_ARTIFACT_TEST=1")     # part of the engine fingerprint
    info = ps.read_aot_model_info(path)
    assert info["engine_matches"] is False
    assert ps.load_aot_model(path).parameter_names() == ["ka", "ke", "v"]


@pytest.mark.gpu
def test_loaded_artifact_launches_the_shipped_device_code(ps, tmp_path, monkeypatch):
    from pharmsol_b200 import _lib
    monkeypatch.setenv("PHARMSOL_B200_CUBIN_CACHE", str(tmp_path / "cache"))
    rng = np.random.default_rng(5)
    subjects = [ps.Subject(f"s{i}", [("bolus", 0.0, 100.0 + i, "po"), ("bolus", 12.0, 50.0, "po")] +
                           [("observation", float(t), float(rng.uniform(0.5, 4.0)), "plasma") for t in (1, 2, 4, 8, 13, 16, 24)]) for i in range(7)]
    data = ps.Data(subjects)
    ems = ps.AssayErrorModels().add("plasma", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0.0, 0.0), 0.0))
    spp = np.column_stack([rng.uniform(0.5, 1.5, 40), rng.uniform(0.05, 0.3, 40), rng.uniform(20, 60, 40)])
    for src, conf, solvers in ((ODE_SRC, lambda e: e.with_solver(ps.OdeSolver.Tsit45).with_tolerances(1e-8, 1e-8), (ps.OdeSolver.Tsit45,)),
                               (ANA_SRC, None, ())):
        monkeypatch.setenv("PHARMSOL_B200_CUBIN_CACHE", str(tmp_path / "cache"))
        jit = ps.Equation.from_dsl(src)
        if conf:
            conf(jit)
        want = ps.log_likelihood_matrix(jit, data, spp, ems)
        path = ps.compile_module_source_to_aot(src, tmp_path / f"{jit.info['name']}.pkm", solvers=solvers, configure=conf)
        # the artifact must not need NVRTC or the cubin cache at load time
        empty = tmp_path / f"empty_cache_{jit.info['name']}"
        monkeypatch.setenv("PHARMSOL_B200_CUBIN_CACHE", str(empty))
        eq = ps.load_aot_model(path)
        assert eq._model.compile(_lib.context(0)) == "artifact"
        got = ps.log_likelihood_matrix(eq, data, spp, ems)
        assert np.array_equal(got, want)
        assert not empty.exists() or os.listdir(empty) == []
    # a solver the artifact does not carry falls back to NVRTC from the embedded source
    eq = ps.load_aot_model(tmp_path / "artifact_ode.pkm").with_solver(ps.OdeSolver.Dopri5)
    assert eq._model.compile(_lib.context(0)) in ("nvrtc", "cubin-cache")
    rt = ps.compile_module_source_to_runtime(ANA_SRC, ps.RuntimeCompilationTarget.CudaAot(tmp_path / "rt.pkm"))
    assert rt._model.compile(_lib.context(0)) == "artifact"
    assert np.array_equal(ps.log_likelihood_matrix(rt, data, spp, ems), want)
