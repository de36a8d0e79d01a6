import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import SHAPES, synth_data, make_sampler, bo
dims4, ntr, nte = SHAPES["madelon"]
train, test = synth_data(dims4, ntr, nte)
s = make_sampler(dims4, train, test)
L = bo.Layout(dims4)
rng = np.random.default_rng(5)
w = (rng.standard_normal(L.w_size)).astype(np.float32); w[L.nbt_index] = 3.0
r = s.eval(0, 0, w=w, grads=True, out=True)
th = w.astype(np.float64)
out, cache = bo.forward(L, th, train.astype(np.float64), update_buffers=False)
g = bo.backward(L, th, cache)
gd = r["grads"].astype(np.float64)
for k in bo.SORTED_KEYS:
    if k in bo.BUFFER_KEYS: continue
    sl = L.slice(k)
    a, b = gd[sl], g[sl]
    err = np.abs(a - b)
    print(k, a.shape, "relerr", np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30), "nbad", int((err > 1e-3 * np.abs(b).max()).sum()))
    if k.endswith("weight") and a.size > 1000:
        shp = {"encode.0.weight": (dims4[1], dims4[0]), "encode.2.weight": (dims4[2], dims4[1]), "encode.4.weight": (dims4[3], dims4[2]),
               "decode.1.weight": (dims4[2], dims4[3]), "decode.3.weight": (dims4[1], dims4[2]), "decode.5.weight": (dims4[0], dims4[1])}.get(k)
        if shp:
            e2 = err.reshape(shp) > 1e-3 * np.abs(b).max()
            rows = np.where(e2.any(1))[0]; cols = np.where(e2.any(0))[0]
            if len(rows): print("   bad rows", rows.min(), rows.max(), len(rows), "bad cols", cols.min(), cols.max(), len(cols))
for k in ("decode.5.bias", "encode.2.bias", "decode.3.bias"):
    sl = L.slice(k)
    print(k, "dev", gd[sl][:6], "ref", g[sl][:6], "ratio", (gd[sl] / g[sl])[:6], (gd[sl] / g[sl])[-6:])
