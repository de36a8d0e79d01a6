"""Shared helpers for the GPU parity tests: run the CUDA sampler and the CPU oracle on identical
weights, proposals and uniforms (the device's Philox draws are read back and fed to the oracle)."""
import numpy as np

from oracle import bae_oracle as bo


def synth_data(dims4, n_train, n_test, seed=1234):
    """SURVEY 8d synthetic inputs: U[0,1) float32 (mimics MinMaxScaler output, MAIN:108,126,133)."""
    train = np.random.default_rng(seed).random((n_train, dims4[0]), dtype=np.float32)
    test = np.random.default_rng(seed + 1).random((n_test, dims4[0]), dtype=np.float32)
    return train, test


def oracle_rnd_from_device(sampler, chain, n_steps, first_step=0):
    lx, noise, en, u = [], [], [], []
    for i in range(first_step, first_step + n_steps):
        d = sampler.debug_draws(chain, i)
        lx.append(d["lx"]); noise.append(d["noise"]); en.append(d["eta_noise"]); u.append(d["u"])
    return dict(lx=np.array(lx), noise=np.array(noise), eta_noise=np.array(en), u=np.array(u))


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
