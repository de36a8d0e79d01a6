"""Freeze a golden vector of the reference's Gelman-Rubin computation (run in the build container only).

    python -m oracle.make_gr_golden          # writes tests/golden/gr_rhat.npz

`Bayesian/Gelman Reuben.py` ("GR") is a top-level script that reads the per-temperature weight files and computes
R-hat in its last 60 lines (GR:676-735). This generator executes exactly those lines of the UNMODIFIED script text
(read from /root/reference at generation time, nothing is copied into the repo) on synthetic weight traces that are
2-decimal quantised like the files the sampler writes (MAIN:641-677 `%1.2f`), and stores inputs and outputs.
Stand-ins: `np.float` (removed from numpy >= 1.24) -> float, `file1.write` -> no-op, `print` -> captures the arrays.
"""
from __future__ import annotations

import os
import types

import numpy as np

GR_PATH = "/root/reference/Bayesian/Gelman Reuben.py"
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gr_rhat.npz")
NAMES = ["weight0", "weight100", "weight2000", "weight3000", "weight4000", "weight5000", "weight6000", "weight7000",
         "weight8000", "weight9000", "weight10000", "weight11000"]   # stacking order of GR:690


def run_reference_block(weights):
    """weights: dict name -> [no_chains, no_samples_b] array (the layout GR fills at GR:7-18, 31-100)."""
    lines = open(GR_PATH).read().split("\n")
    first = next(i for i, l in enumerate(lines) if l.strip() == "weight0 = weight0.T")
    block = "\n".join(lines[first:])
    shim = types.SimpleNamespace(**{k: getattr(np, k) for k in ("stack", "array", "cov")})
    shim.float = float
    captured = []

    def fake_print(*a):
        captured.append(np.array(a[0], dtype=np.float64).copy())

    ns = dict(np=shim, file1=types.SimpleNamespace(write=lambda s: None), print=fake_print)
    ns.update({k: np.array(v, dtype=np.float64) for k, v in weights.items()})
    exec(compile(block, GR_PATH, "exec"), ns)
    return dict(simple=captured[0], advanced=captured[1], data_shape=np.array(ns["data"].shape))


def main():
    rng = np.random.default_rng(20260)
    chains, samples = 8, 500
    weights = {}
    for k, name in enumerate(NAMES):
        # AR(1) traces with chain-dependent offsets: R-hat visibly above 1 for some parameters
        x = np.zeros((chains, samples))
        e = rng.standard_normal((chains, samples))
        for t in range(1, samples):
            x[:, t] = 0.9 * x[:, t - 1] + 0.1 * e[:, t]
        x += (k % 4) * 0.05 * rng.standard_normal((chains, 1)) + k * 0.1
        weights[name] = np.round(x, 2)
    out = run_reference_block(weights)
    np.savez_compressed(GOLDEN, names=np.array(NAMES), simple=out["simple"], advanced=out["advanced"],
                        data_shape=out["data_shape"], **{"w_" + k: v for k, v in weights.items()})
    print("wrote", GOLDEN, "simple", out["simple"][:4], "advanced", out["advanced"][:4], "shape", out["data_shape"])


if __name__ == "__main__":
    main()
