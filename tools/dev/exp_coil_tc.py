"""Experiment: Coil-2000 widths on the tcgen05 engine (BAE_TC_MIN_WIDTH=48): forward/backward parity + speed."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bae_oracle as bo
from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
from tests.parity_util import synth_data, rel
dims4, ntr, nte = (85, 70, 60, 50), 4366, 1456
train, test = synth_data(dims4, ntr, nte)
L = bo.Layout(dims4)
for gp in (2, 1):
    s = DeviceSampler(dims4, train, test, 1, 1, 4, [1.0], gemm_path=gp)
    rng = np.random.default_rng(5)
    for scale in (1.0, 0.3):
        w = (rng.standard_normal(L.w_size) * scale).astype(np.float32); w[L.nbt_index] = 3.0
        r = s.eval(0, 0, w=w, grads=True)
        th = w.astype(np.float64)
        out, c = bo.forward(L, th, train.astype(np.float64), update_buffers=False)
        g = bo.backward(L, th, c)
        sse = float(np.sum((out - train) ** 2))
        print("path", gp, "scale", scale, "sse rel %.2e" % (abs(r["sse"] - sse) / sse), "grad rel %.2e" % rel(r["grads"], g), flush=True)
    s.close()
for gp in (2, 1):
    for C in (8, 64):
        T = np.tile(2.0 ** (np.arange(8) / 7), C // 8)
        s = DeviceSampler(dims4, train, test, C, 8, 400, T, swap_interval=5, gemm_path=gp)
        rng = np.random.default_rng(0)
        for j in range(C):
            s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
        for _ in range(3):
            s.run(5)
        s.synchronize(); t = time.time()
        for _ in range(10):
            s.run(5)
        s.synchronize(); dt = time.time() - t
        print("path", gp, "chains", C, "chain-steps/s %.0f" % (C * 50 / dt), "ms/round %.2f" % (dt * 100), flush=True)
        s.close()
