"""Glue between ``benches/workloads.py`` (pure data) and the two implementations: the product
(``pharmsol_b200``) and the CPU oracle (``oracle``; imported lazily and only by tests, smoke and the
cpu_baseline / --impl reference legs of bench.py)."""
from __future__ import annotations

import numpy as np

OUTPUT_ORDER = {"c4": ["cp", "effect"]}
ERRKIND = {"additive": 1, "proportional": 2}


def product_objects(w, device=None):
    import pharmsol_b200 as ps
    eq = ps.Equation.from_dsl(w["dsl"], device)
    data = ps.Data([ps.Subject(i, ops) for i, ops in w["subjects"]])
    ems = ps.AssayErrorModels()
    for label, (kind, factor, poly) in w["error_models"].items():
        model = (ps.AssayErrorModel.additive if kind == "additive" else ps.AssayErrorModel.proportional)(ps.ErrorPoly(*poly), factor)
        ems.add(label, model)
    return eq, data, ems


def oracle_objects(w, **model_kw):
    import oracle as O
    model = O.Model(w["oracle_model"], **model_kw)
    data = O.Data([O.Subject(ops, i) for i, ops in w["subjects"]])
    outputs = OUTPUT_ORDER.get(w["name"], list(w["error_models"].keys()))
    ems = O.ErrorModels([w["error_models"].get(o) for o in outputs])
    return model, data, ems


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
