"""tests/golden/sde_c5_oracle.json — seed-averaged log-likelihoods of the RESTATED reference SDE path (oracle: adaptive
Euler-Maruyama as written in sde/em.rs:134-167 + mean-prediction likelihood / particle filter, sde/mod.rs:387-433,
526-577, 747-767) on slices of the C5 workload, for the statistical parity tests of the device path
(tests/test_gpu_sde_parity.py).  The reference draws from an unseeded thread-local ChaCha stream, so there is no stream
to match: parity is |mean_device - mean_oracle| <= 3 SE over >= 64 seeds.

    python scripts/gen_sde_golden.py          # ~10 min on 8 cores
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from benches import harness as H, workloads as W  # noqa: E402

CASES = [
    # name, nsub, nspp, particles, seeds: 64 pairs spanning the C5 parameter box (ke, sigma, v) at 128 particles ...
    dict(name="box64", nsub=8, nspp=8, particles=128, nseed=64),
    # ... and a slice of the BASELINE-size population at its own particle count (1,000)
    dict(name="baseline_slice", nsub=4, nspp=6, particles=1000, nseed=64),
]


def main():
    out = {"generator": "scripts/gen_sde_golden.py", "oracle": "oracle/ (restated EM + PF)", "seed0": 770000, "cases": []}
    for c in CASES:
        w = W.make("c5", nsub=c["nsub"], nspp=c["nspp"], particles=c["particles"])
        om, od, oe = H.oracle_objects(w, particles=c["particles"])
        rec = dict(c)
        for mode, key in ((0, "mean_prediction"), (1, "particle_filter")):
            t0 = time.time()
            o = np.stack([om.log_likelihood_matrix(od, w["support_points"], oe, seed=out["seed0"] + 1000 * mode + s, sde_mode=mode) for s in range(c["nseed"])])
            finite = np.isfinite(o).all(axis=0)
            rec[key] = {"mean": np.where(finite, o.mean(axis=0), np.nan).tolist(), "var": np.where(finite, o.var(axis=0, ddof=1), np.nan).tolist(),
                        "all_finite": finite.tolist()}
            print(c["name"], key, f"{time.time() - t0:.1f} s", flush=True)
        out["cases"].append(rec)
    path = os.path.join(ROOT, "tests", "golden", "sde_c5_oracle.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path)


if __name__ == "__main__":
    main()
