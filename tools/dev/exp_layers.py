"""Per-layer accuracy of one forward/backward (GPU): device activations / deltas / gradients against the float64 oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bae_oracle as bo
from bayesianautoencoders_b200 import DeviceSampler
from tests.parity_util import synth_data, rel
dims4, ntr, nte = (500, 450, 400, 300), 2000, 1800
train, test = synth_data(dims4, ntr, nte)
gp = int(os.environ.get("EXP_GEMM_PATH", "2"))
s = DeviceSampler(dims4, train, test, 1, 1, 4, [1.0], gemm_path=gp)
L = bo.Layout(dims4)
rng = np.random.default_rng(5)
for scale in (1.0, 0.3):
    w = (rng.standard_normal(L.w_size) * scale).astype(np.float32)
    w[L.nbt_index] = 3.0
    r = s.eval(0, 0, w=w, grads=True)
    th = w.astype(np.float64)
    out, c = bo.forward(L, th, train.astype(np.float64), update_buffers=False)
    g = bo.backward(L, th, c)
    d = {}
    for name, ref, dd in (("H1", c["h1"], 450), ("H2", c["h2"], 400), ("Y", c["y"], 300), ("H4", c["h4"], 400), ("H5", c["h5"], 450)):
        b = s.debug_buffer(name, 1, ntr)[:, :dd]
        d[name] = rel(b, ref)
    z = s.debug_buffer("Z", 1, ntr)[:, :300]
    zo = c["zhat"] / c["inv"] + c["mu"]
    d["Z"] = rel(z, zo)
    d["Z-mean"] = rel(z - z.mean(0), zo - zo.mean(0))
    d["|mean|/std"] = float(np.median(np.abs(zo.mean(0)) / zo.std(0)))
    print("path", gp, "scale", scale, "sse rel", abs(r["sse"] - np.sum((out - train) ** 2)) / np.sum((out - train) ** 2))
    print("  fwd ", {k: "%.2e" % v for k, v in d.items()})
    gd = r["grads"].astype(np.float64)
    print("  grad whole %.2e" % rel(gd, g), {k.replace("code.", ""): "%.1e" % rel(gd[L.slice(k)], g[L.slice(k)]) for k in bo.SORTED_KEYS if k not in bo.BUFFER_KEYS and k != "encode.4.bias"})
s.close()
