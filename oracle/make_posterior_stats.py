"""Posterior statistics of the CPU oracle at the Swiss Roll shape (BASELINE.json configs[0]: 8 replicas, 8000 samples),
frozen as the fixture of the GPU acceptance test `tests/test_gpu_posterior.py` (TEST INFRASTRUCTURE).

    python -m oracle.make_posterior_stats          # writes tests/golden/swiss_posterior_stats.npz  (~2 min on 8 cores)

R independent parallel-tempering runs (own numpy draws, own torch-default-style initial models), each 8 replicas x
1000 iterations with an exchange sweep every 5, through `bae_oracle.run_pt` (MAIN:1013-1072). Stored per run and replica:
post-burn-in means of the recorded MSE traces (MAIN:542-560) and of the proposals' MSE, acceptance percentage
(MAIN:694-696), and per run the swap percentage (MAIN:1084) and the Gelman-Rubin R-hat of weight[0] / weight[100]
(the traced entries that exist at this shape, GR:690-702).
"""
from __future__ import annotations

import multiprocessing as mp
import os

import numpy as np

DIMS4, NTR, NTE, N, S, SI, RUNS = (3, 10, 5, 2), 3750, 1250, 8, 1000, 5, 8
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "swiss_posterior_stats.npz")


def _theta_init(dims4, rng):
    """torch-default-style initial state_dict, flat sorted-key order (what `Model()` holds, MAIN:270)."""
    from .bae_oracle import Layout
    L = Layout(dims4)
    th = np.zeros(L.w_size)
    for name in ("encode.0", "encode.2", "encode.4", "decode.1", "decode.3", "decode.5"):
        fan_in = L.shapes[name + ".weight"][1]
        b = 1.0 / np.sqrt(fan_in)
        th[L.slice(name + ".weight")] = rng.uniform(-b, b, L.numel(name + ".weight")).astype(np.float32)
        th[L.slice(name + ".bias")] = rng.uniform(-b, b, L.numel(name + ".bias")).astype(np.float32)
    th[L.slice("decode.0.weight")] = 1.0
    th[L.slice("decode.0.running_var")] = 1.0
    return th


def one_run(seed):
    from threadpoolctl import threadpool_limits
    from . import bae_oracle as bo
    train = np.random.default_rng(1234).random((NTR, DIMS4[0]), dtype=np.float32)
    test = np.random.default_rng(1235).random((NTE, DIMS4[0]), dtype=np.float32)
    L = bo.Layout(DIMS4)
    temps = bo.temperature_ladder(N, 2)
    rng = np.random.default_rng([777, seed])
    th0 = [_theta_init(DIMS4, rng) for _ in range(N)]
    rnd = [dict(w0_randn=rng.standard_normal(L.w_size), lx=rng.random(S),
                noise=(rng.standard_normal((S, L.w_size)) * 0.005).astype(np.float32),
                eta_noise=rng.standard_normal(S) * 0.2, u=rng.random(S)) for _ in range(N)]
    swap_u = list(rng.random((S // SI) * (N - 1)))
    w = [[] for _ in range(N)]

    def on_step(j, i, rep, rec):
        w[j].append(rep.traced_weights()[:2])

    with threadpool_limits(limits=1):
        reps, records, swap_log = bo.run_pt(DIMS4, train, test, temps, S, SI, th0, rnd, swap_u=swap_u, on_step=on_step)
    b = S // 2
    rec_tr, rec_te = np.zeros((N, S)), np.zeros((N, S))
    for j in range(N):
        a = c = 0.0
        for i, r in enumerate(records[j]):
            if r.accepted:
                a, c = r.mse_train, r.mse_test
            rec_tr[j, i], rec_te[j, i] = a, c
    wt = np.array(w)[:, b:, :]                                   # [chains, samples, 2]
    B_on_n = wt.mean(axis=1).var(axis=0)
    W = wt.var(axis=1).mean(axis=0)
    ns = wt.shape[1]
    rhat = ((ns / (ns - 1)) * W + B_on_n + B_on_n / N) / W       # GR:699-702 (textbook orientation)
    return dict(
        mse_train=rec_tr[:, b:].mean(1), mse_test=rec_te[:, b:].mean(1),
        prop_mse_train=np.array([np.mean([r.mse_train for r in records[j][b:]]) for j in range(N)]),
        prop_mse_test=np.array([np.mean([r.mse_test for r in records[j][b:]]) for j in range(N)]),
        accept=np.array([100.0 * sum(r.accepted for r in records[j]) / S for j in range(N)]),
        langevin=np.array([100.0 * sum(r.langevin for r in records[j]) / S for j in range(N)]),
        swap=100.0 * sum(1 for s in swap_log if s[2]) / len(swap_log), rhat=rhat)


def main():
    with mp.get_context("fork").Pool(min(RUNS, os.cpu_count() or 1)) as pool:
        res = pool.map(one_run, range(RUNS))
    out = {k: np.array([r[k] for r in res]) for k in res[0]}
    np.savez_compressed(GOLDEN, dims4=np.array(DIMS4), n_train=NTR, n_test=NTE, n_chains=N, samples=S, swap_interval=SI, **out)
    for k, v in out.items():
        print(k, np.round(np.mean(v, axis=0), 4))


if __name__ == "__main__":
    main()
