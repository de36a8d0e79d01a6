import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import bae_oracle as bo
from bayesianautoencoders_b200 import DeviceSampler
from tests.parity_util import synth_data, rel
shapes = {"m256": ((256, 256, 256, 256), 512, 256), "madelon": ((500, 450, 400, 300), 2000, 1800)}
which = sys.argv[1:] or ["m256", "madelon"]
for name in which:
    dims4, ntr, nte = shapes[name]
    train, test = synth_data(dims4, ntr, nte)
    L = bo.Layout(dims4)
    w = (np.random.default_rng(5).standard_normal(L.w_size) * 0.3).astype(np.float32); w[L.nbt_index] = 3
    th = w.astype(np.float64)
    out, cache = bo.forward(L, th, train.astype(np.float64), update_buffers=False)
    g = bo.backward(L, th, cache)
    sse = float(np.sum((out - train) ** 2))
    for path in (2,):
        s = DeviceSampler(dims4, train, test, 1, 1, 20, [1.0], gemm_path=path)
        t = time.time(); r = s.eval(0, 0, w=w, grads=True, out=True); dt = time.time() - t
        print(name, 'path', s.gemm_path(), 'eval s %.4f' % dt, 'sse', r['sse'], sse, 'rel %.2e' % ((r['sse'] - sse) / sse), 'out rel %.2e' % rel(r['out'], out),
              'grad rel %.2e' % rel(r['grads'], g), flush=True)
        for k in bo.SORTED_KEYS:
            if k in bo.BUFFER_KEYS: continue
            sl = L.slice(k); print('     %-16s %.2e  dev max %.3e ref max %.3e nan %d' % (k, rel(r['grads'][sl], g[sl]), np.abs(r['grads'][sl]).max(), np.abs(g[sl]).max(), int(np.isnan(r['grads'][sl]).sum())))
        s.close()
