#!/usr/bin/env python3
"""Small driver for profiling: N Madelon-shape chains, a few PT rounds (used under ncu)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianautoencoders_b200 import DeviceSampler, default_theta_init  # noqa: E402
from bench import SHAPES, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="madelon")
ap.add_argument("--chains", type=int, default=64)
ap.add_argument("--rounds", type=int, default=1)
ap.add_argument("--steps-per-round", type=int, default=5)
ap.add_argument("--gemm-path", type=int, default=0)
ap.add_argument("--profile-round", type=int, default=-1, help="cudaProfilerStart/Stop around this round (ncu --profile-from-start off)")
a = ap.parse_args()
dims4, ntr, nte = SHAPES[a.shape]
train, test = synth(dims4, ntr, nte)
rung = 8
T = np.tile(2.0 ** (np.arange(rung) / (rung - 1)), a.chains // rung)
s = DeviceSampler(dims4, train, test, a.chains, rung, 1000, T, swap_interval=5, gemm_path=a.gemm_path)
rng = np.random.default_rng(0)
for j in range(a.chains):
    s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
for r in range(a.rounds):
    if r == a.profile_round:
        s.synchronize(); s.lib.bae_debug_profiler(1)
    s.run(a.steps_per_round)
    s.synchronize()
    if r == a.profile_round:
        s.lib.bae_debug_profiler(0)
    print("round", r, "ms", s.last_ms(), flush=True)
tr = s.traces(0, 0, a.rounds * a.steps_per_round)
print("accepted", tr["accepted"].tolist(), "langevin", tr["langevin"].tolist())
s.close()
