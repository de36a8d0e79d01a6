"""Forward / backward of one Madelon-shape chain on the device (3xFP16 or 3xTF32 engine) against the oracle, layer by layer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesianautoencoders_b200 import DeviceSampler
from oracle import bae_oracle as bo
from tests.parity_util import rel, synth_data

dims4, ntr, nte = (500, 450, 400, 300), 2000, 1800
if len(sys.argv) > 1 and sys.argv[1] == "small":
    dims4, ntr, nte = (500, 450, 400, 300), 600, 300
train, test = synth_data(dims4, ntr, nte)
s = DeviceSampler(dims4, train, test, 1, 1, 10, np.ones(1), seed=7, gemm_path=2)
L = bo.Layout(dims4)
rng = np.random.default_rng(5)
for scale in (1.0, 0.3):
    w = (rng.standard_normal(L.w_size) * scale).astype(np.float32)
    w[L.nbt_index] = 3.0
    r = s.eval(0, 0, w=w, grads=True, out=True)
    th = w.astype(np.float64)
    out, cache = bo.forward(L, th, train.astype(np.float64), update_buffers=False)
    g = bo.backward(L, th, cache)
    sse = float(np.sum((out - train) ** 2))
    print("scale", scale, "sse rel", abs(r["sse"] - sse) / sse, "out rel", rel(r["out"], out))
    gd = r["grads"].astype(np.float64)
    print("  grad rel", np.linalg.norm(gd - g) / np.linalg.norm(g))
    for k in bo.SORTED_KEYS:
        if k in bo.BUFFER_KEYS:
            continue
        sl = L.slice(k)
        print("   %-16s %.3e" % (k, np.linalg.norm(gd[sl] - g[sl]) / max(np.linalg.norm(g[sl]), 1e-30)))
    ref = dict(H1=cache["h1"], H2=cache["h2"], Z=cache["zhat"] / cache["inv"] + cache["mu"], Y=cache["y"], H4=cache["h4"], H5=cache["h5"])
    for name in ("H1", "H2", "Z", "Y", "H4", "H5"):
        try:
            buf = s.debug_buffer(name, 0, ntr)     # the probe ran alone: slot 0 of the activation buffers
        except Exception as e:
            print("   no buffer", name, e); break
        d = ref[name].shape[1]
        print("   %-3s rel %.3e   max|dev| %.3e" % (name, rel(buf[:, :d], ref[name]), np.abs(buf[:, :d]).max()))
        if name == "H1" and scale == 1.0:
            e = np.abs(buf[:, :d] - ref[name])
            print("      err by 64-row block x 30-col block (max):")
            for r0 in range(0, min(ntr, 512), 64):
                print("      ", " ".join("%7.1e" % e[r0:r0 + 64, c0:c0 + 30].max() for c0 in range(0, d, 30)))
s.close()
