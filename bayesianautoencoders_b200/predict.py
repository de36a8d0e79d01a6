"""Posterior-predictive / encoder-feature pass, batched over posterior samples.

The reference runs this once, on the final weights of one replica, at the end of `ptReplica.run()` (MAIN:709-818):
`cae.encode(X)` feeds the downstream classifiers (MAIN:714, 787) and `cae.forward(X)` gives the reconstruction
(MAIN:752). The model is never put in eval mode there, so BatchNorm uses the statistics of the batch it is given.
Here the same pass runs for K weight vectors at once as a chain-batched forward on the device (`bae_forward_all`).
"""
from __future__ import annotations

import numpy as np

from .sampler import DeviceSampler


class PosteriorPredictive:
    """K posterior samples (flat `getparameters` vectors, MAIN:228-241) evaluated on one data matrix.

        pp = PosteriorPredictive(dims4, X, K)
        pp.load(k, w_k) for every sample        (or pp.load_all(W[K, w_size]))
        feats = pp.encode()       # [K, rows, enc_shape]   cae.encode(X)
        recon = pp.reconstruct()  # [K, rows, in_shape]    cae.forward(X)
        mse   = pp.mse()          # [K]                    reconstruction MSE per sample
    """

    def __init__(self, dims4, data, n_samples, device=0, gemm_path=0):
        data = np.ascontiguousarray(data, dtype=np.float32)
        if data.ndim != 2 or data.shape[1] != int(dims4[0]) or data.shape[0] < 2:
            raise ValueError("data must be a [rows >= 2, in_shape] matrix (BatchNorm needs batch statistics)")
        self.K = int(n_samples)
        self.rows = data.shape[0]
        self.dims4 = tuple(int(x) for x in dims4)
        # one chain slot per posterior sample; the sampler's step program is never run on this handle
        self.s = DeviceSampler(self.dims4, data, data, self.K, 1, 1, np.ones(self.K), device=device, gemm_path=gemm_path)
        self.w_size = self.s.w_size
        self._loaded = np.zeros(self.K, bool)

    def load(self, k, w):
        w = np.ascontiguousarray(w, dtype=np.float32)
        if w.shape != (self.w_size,):
            raise ValueError(f"expected a flat vector of {self.w_size} entries in getparameters order")
        self.s.set_state(int(k), theta=w)
        self._loaded[k] = True

    def load_all(self, W):
        W = np.asarray(W)
        for k in range(self.K):
            self.load(k, W[k])

    def _run(self, features, recon):
        if not self._loaded.all():
            raise RuntimeError("load every posterior sample first")
        return self.s.forward_all(0, features=features, recon=recon)

    def encode(self):
        return self._run(True, False)["features"]

    def reconstruct(self):
        return self._run(False, True)["recon"]

    def mse(self):
        return self._run(False, False)["sse"] / (self.rows * self.dims4[0])

    def close(self):
        self.s.close()
