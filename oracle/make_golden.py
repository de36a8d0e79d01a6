"""Freeze golden vectors from the UNMODIFIED reference (run in the build container only).

    python -m oracle.make_golden            # writes tests/golden/*.npz

Each case draws its randomness from a fixed numpy Generator, feeds it to the reference through
`oracle/ref_driver.py`, and stores: the torch-default initial state_dict (flat, sorted-key), the
injected draws, the full-precision per-step traces the reference would have written with
np.savetxt (MAIN:631-692), every gradient produced by `langevin_gradient` (MAIN:224), and the swap
log of `swap_procedure` (MAIN:962-1011). The files are small (< 300 KB total) and committed.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import ref_driver
from .bae_oracle import Layout

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: dims4, n_train, n_test, temperatures, samples/chain, swap_interval, do_swaps, seed
    "swiss_1chain": dict(dims4=(3, 10, 5, 2), n_train=300, n_test=100, temps=[1.5], S=12, si=5, swaps=False, seed=11),
    "swiss_pt4": dict(dims4=(3, 10, 5, 2), n_train=256, n_test=64, temps=None, n_chains=4, S=20, si=5, swaps=True, seed=12),
    "mid_pt3": dict(dims4=(24, 20, 16, 12), n_train=96, n_test=40, temps=None, n_chains=3, S=10, si=5, swaps=True, seed=13),
    # experiment_type = 1 (MAIN:72): Adam while i <= pt = 6, stateless SGD(lr = 0.1) afterwards (MAIN:454-470)
    "mid_sgd2": dict(dims4=(24, 20, 16, 12), n_train=96, n_test=40, temps=None, n_chains=2, S=12, si=5, swaps=True, seed=14,
                     experiment_type=1),
}


def make_randomness(L: Layout, S: int, n_chains: int, seed: int, step_w=0.005, step_eta=0.2):
    out = []
    for j in range(n_chains):
        r = np.random.default_rng([seed, j])
        out.append(dict(
            w0_randn=r.standard_normal(L.w_size),
            lx=r.random(S),
            noise=(r.standard_normal((S, L.w_size)) * step_w).astype(np.float32),
            eta_noise=r.standard_normal(S) * step_eta,
            u=r.random(S)))
    return out


def _find(saved, prefix, exclude=None):
    for k in saved:
        if k.startswith(prefix) and (exclude is None or exclude not in k):
            return saved[k]
    raise KeyError(prefix)


def build_case(name, spec):
    from .bae_oracle import temperature_ladder
    L = Layout(spec["dims4"])
    rng = np.random.default_rng(spec["seed"])
    train = rng.random((spec["n_train"], spec["dims4"][0]), dtype=np.float32)
    test = rng.random((spec["n_test"], spec["dims4"][0]), dtype=np.float32)
    temps = spec["temps"] if spec["temps"] is not None else list(temperature_ladder(spec["n_chains"], 2))
    n = len(temps)
    rnd = make_randomness(L, spec["S"], n, spec["seed"])
    n_rounds = spec["S"] // spec["si"]
    swap_u = list(np.random.default_rng([spec["seed"], 999]).random(n_rounds * max(n - 1, 1)))
    res = ref_driver.run_reference_pt(spec["dims4"], train, test, temps, spec["S"], spec["si"], rnd,
                                      swap_u=swap_u, do_swaps=spec["swaps"], torch_seed=spec["seed"],
                                      experiment_type=spec.get("experiment_type", 2))
    out = dict(dims4=np.array(spec["dims4"]), train=train, test=test, temps=np.array(temps, dtype=np.float64),
               S=spec["S"], swap_interval=spec["si"], do_swaps=spec["swaps"], swap_u=np.array(swap_u),
               swap_log=np.array(res["swap_log"], dtype=np.int64).reshape(-1, 3),
               num_swap=res["num_swap"], total_swap_proposals=res["total_swap_proposals"],
               experiment_type=spec.get("experiment_type", 2))
    for j, c in enumerate(res["ctxs"]):
        s = c.saved
        out[f"c{j}_theta_init"] = c.theta_init
        for k, v in rnd[j].items():
            out[f"c{j}_rnd_{k}"] = np.asarray(v)
        out[f"c{j}_sum_value"] = _find(s, "sum_value_")
        out[f"c{j}_likelihood"] = _find(s, "likelihood_value_", exclude="proposal")
        out[f"c{j}_likelihood_proposal"] = _find(s, "likelihood_value_proposal")
        out[f"c{j}_mse_train"] = _find(s, "mse_train_chain_")
        out[f"c{j}_mse_test"] = _find(s, "mse_test_chain_")
        out[f"c{j}_acc_train"] = _find(s, "acc_train_chain_")
        out[f"c{j}_u_value"] = _find(s, "u_value_")
        out[f"c{j}_weight0"] = _find(s, "weight[0]_")
        out[f"c{j}_weight100"] = _find(s, "weight[100]_")
        out[f"c{j}_grads"] = np.array(c.grads)
        out[f"c{j}_lik_calls"] = np.array(c.lik_calls)
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        out = build_case(name, spec)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
