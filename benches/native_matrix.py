"""The reference's own criterion shapes (benches/native_matrix.rs:23-24, benches/common/mod.rs:117-271):
`native/likelihood-matrix` at 32 subjects x 64 support points, workloads `1cpt-12h-po` (1 bolus, 9 obs) and
`2cpt-120h-q12h` (10 boluses, 14 obs), analytical and ODE (rtol = atol = 1e-4 as in the reference bench), additive
ErrorPoly(0.1, 0.1, 0, 0).  Prints one JSON line per case: GPU time per matrix (host-buffer call, median of 200) and
the restated CPU oracle on all host cores, so the numbers are comparable with `cargo bench --bench native_matrix`.
Tiny matrices (2,048 pairs) are launch-latency bound on a GPU — that is the point of publishing them."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

import pharmsol_b200 as ps
from benches import workloads as W

DSL = {
    ("short", "analytical"): "name = bench_short_a\nkind = analytical\nparams = ka, ke, v\nstates = gut, central\noutputs = plasma\nbolus(po) -> gut\n"
                             "structure = one_compartment_with_absorption\nout(plasma) = central / v ~ continuous()\n",
    ("short", "ode"): "name = bench_short_o\nkind = ode\nparams = ka, ke, v\nstates = gut, central\noutputs = plasma\nbolus(po) -> gut\n"
                      "dx(gut) = -ka * gut\ndx(central) = ka * gut - ke * central\nout(plasma) = central / v ~ continuous()\n",
    ("repeat", "analytical"): "name = bench_repeat_a\nkind = analytical\nparams = ke, kcp, kpc, v\nstates = central, peripheral\noutputs = plasma\nbolus(iv) -> central\n"
                              "structure = two_compartments\nout(plasma) = central / v ~ continuous()\n",
    ("repeat", "ode"): "name = bench_repeat_o\nkind = ode\nparams = ke, kcp, kpc, v\nstates = central, peripheral\noutputs = plasma\nbolus(iv) -> central\n"
                       "dx(central) = -(ke + kcp) * central + kpc * peripheral\ndx(peripheral) = kcp * central - kpc * peripheral\n"
                       "out(plasma) = central / v ~ continuous()\n",
}


def main():
    import oracle as O
    for (workload, family), src in DSL.items():
        w = W.reference_bench(workload)
        eq = ps.Equation.from_dsl(src)
        if family == "ode":
            eq.with_solver(ps.OdeSolver.Dopri5).with_tolerances(1e-4, 1e-4)
        data = ps.Data([ps.Subject(i, o) for i, o in w["subjects"]])
        ems = ps.AssayErrorModels().add("plasma", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0.0, 0.0), 0.0))
        spp = w["support_points"]
        for _ in range(20):
            psi = ps.log_likelihood_matrix(eq, data, spp, ems)
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            psi = ps.log_likelihood_matrix(eq, data, spp, ems)
            ts.append(time.perf_counter() - t0)
        gpu_s = float(np.median(ts))
        one = np.ascontiguousarray(spp[:1])       # an optimiser's cost function: every subject under ONE support point
        t1 = []
        for _ in range(220):
            t0 = time.perf_counter()
            ps.log_likelihood_matrix(eq, data, one, ems)
            t1.append(time.perf_counter() - t0)
        one_s = float(np.median(t1[20:]))
        om = O.Model(f"bench_{workload}_{family}", **(dict(solver="dopri5", rtol=1e-4, atol=1e-4) if family == "ode" else {}))
        od = O.Data([O.Subject(o, i) for i, o in w["subjects"]])
        oe = O.ErrorModels([w["error_models"]["plasma"]])
        cs = []
        for _ in range(30):
            t0 = time.perf_counter()
            ref = om.log_likelihood_matrix(od, spp, oe)
            cs.append(time.perf_counter() - t0)
        cpu_s = float(np.median(cs))
        err = float(np.max(np.abs(psi - ref) / (np.abs(ref) + 14)))
        print(json.dumps({"bench": f"native/likelihood-matrix/{'1cpt-12h-po' if workload == 'short' else '2cpt-120h-q12h'}/{family}", "nsub": 32, "nspp": 64,
                          "gpu_us_per_matrix": gpu_s * 1e6, "gpu_us_per_single_column": one_s * 1e6, "gpu_pairs_per_s": 2048 / gpu_s, "cpu_oracle_us_per_matrix": cpu_s * 1e6,
                          "cpu_oracle_pairs_per_s": 2048 / cpu_s, "cpu_threads": os.cpu_count(), "max_scaled_ll_diff": err}), flush=True)


if __name__ == "__main__":
    main()
