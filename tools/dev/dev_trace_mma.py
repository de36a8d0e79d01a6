"""MMA-issue timeline of CTA pair 0 in one GEMM phase (BAE_TC_TRACE=1, BAE_TC_TRACE_PHASE=<phase index>): per tile (or K segment)
the number of K blocks, the in-tile pace and the gap to the next tile."""
import os, sys, numpy as np
os.environ["BAE_TC_TRACE"] = "1"
ph = os.environ.setdefault("BAE_TC_TRACE_PHASE", "2")
sys.path.insert(0, '.')
from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
from bayesianautoencoders_b200 import _native as N
from bench import SHAPES, synth
dims4, ntr, nte = SHAPES["madelon"]
train, test = synth(dims4, ntr, nte)
C = int(os.environ.get('TRACE_CHAINS', '64'))
s = DeviceSampler(dims4, train, test, C, 8, 100, np.tile(2.0 ** (np.arange(8) / 7), C // 8))
rng = np.random.default_rng(0)
for j in range(C): s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
buf = np.zeros(4096, np.int64)
s.step(1); s.synchronize()
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 1))
s.step(1); s.synchronize()
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 0))
ev = buf[1024:2048].reshape(-1, 2)
ev = [(int(a), int(b)) for a, b in ev if b != 0]
tiles, cur = [], []
for a, b in ev:
    if a == 1000:
        tiles.append(cur); cur = []
    else:
        cur.append((a, b))
if cur: tiles.append(cur)
t0 = ev[0][1]
print("phase", ph, "tiles", len(tiles), "total cycles", ev[-1][1] - t0)
prev_end = None
for t in tiles:
    ts = np.array([b for _, b in t])
    d = np.diff(ts)
    gap = (ts[0] - prev_end) if prev_end is not None else 0
    print(f"  blocks {len(t):4d}  span {ts[-1] - ts[0]:7d}  pace median {np.median(d) if len(d) else 0:6.0f} mean {d.mean() if len(d) else 0:6.0f}  gap before {gap:6d}")
    prev_end = ts[-1]
s.close()
