"""examples/native_matrix.c — the C ABI driven from plain C (the shape of a cgo / Rust-FFI binding): the reference's own
criterion cases (benches/native_matrix.rs:23-24, 32 x 64) built through the Subject builder entry points."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu


@pytest.fixture(scope="module")
def exe(libpath):
    import __graft_entry__ as g
    return g.build_c_caller()


def test_c_caller_builds_warning_free_and_fails_loudly_without_a_device(exe):
    assert os.access(exe, os.X_OK)
    needed = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libpharmsol_cuda.so" in needed and "libtorch" not in needed and "libpython" not in needed
    if not has_gpu():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CUDA device" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_c_caller_matches_the_python_binding(ps, exe):
    import sys
    sys.path.insert(0, ROOT)
    from benches import native_matrix as NM, workloads as W
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = {j["bench"]: j for j in map(json.loads, r.stdout.strip().splitlines())}
    assert len(rows) == 4
    for (workload, family), src in NM.DSL.items():
        w = W.reference_bench(workload)
        eq = ps.Equation.from_dsl(src)
        if family == "ode":
            eq.with_solver(ps.OdeSolver.Dopri5).with_tolerances(1e-4, 1e-4)
        data = ps.Data([ps.Subject(i, o) for i, o in w["subjects"]])
        ems = ps.AssayErrorModels().add("plasma", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0.0, 0.0), 0.0))
        psi = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
        row = rows[f"native/likelihood-matrix/{'1cpt-12h-po' if workload == 'short' else '2cpt-120h-q12h'}/{family}"]
        assert row["first_error_code"] == 0
        # same library, same inputs (the C program rebuilds them independently): bit-identical results
        assert row["psi_00"] == psi[0, 0]
        assert row["psi_sum"] == pytest.approx(float(np.sum(psi.ravel(order="F"))), rel=1e-13)
        assert 5.0 < row["us_per_matrix"] < 5000.0


@pytest.mark.gpu
def test_c_caller_multi_device_context(exe):
    """SURVEY §8b `ctx_create(device_ids, n_dev)` from plain C: two column shards on GPU 0 (every visible GPU when there
    are several) give the single-device matrix bit for bit, and the replicated call returns one matrix per device."""
    import torch
    n = torch.cuda.device_count()
    devs = ",".join(str(k) for k in range(n)) if n > 1 else "0,0"
    r = subprocess.run([exe, "--devices", devs, "4096"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    row = json.loads(r.stdout.strip().splitlines()[-1])
    assert row["matches_single_device"] is True and row["replicated_ptrs"] is True and row["first_error_code"] == 0
    assert row["devices"] == max(n, 2)
