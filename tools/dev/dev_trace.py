"""Timeline of one GEMM phase of CTA 0 (debug): BAE_TC_TRACE=1 + per-phase launches; the trace buffer holds the LAST
GEMM phase of the step (test forward, layer 6). Prints the K-block pace per role and the epilogue stamps."""
import os, sys, numpy as np
os.environ["BAE_TC_TRACE"] = "1"; os.environ["BAE_PHASE_LAUNCH"] = "1"
sys.path.insert(0, '.')
from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
from bayesianautoencoders_b200 import _native as N
from bench import SHAPES, synth
dims4, ntr, nte = SHAPES["madelon"]
train, test = synth(dims4, ntr, nte)
C = int(os.environ.get('TRACE_CHAINS', '16'))
s = DeviceSampler(dims4, train, test, C, 8, 100, np.tile(2.0 ** (np.arange(8) / 7), C // 8))
rng = np.random.default_rng(0)
for j in range(C): s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
buf = np.zeros(4096, np.int64)
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 1))
s.step(1); s.synchronize()
print("step ms", s.last_ms())
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 0))
t0 = None
for role, name in enumerate(["producer", "mma", "convert", "epilogue"]):
    ev = buf[role * 1024:(role + 1) * 1024].reshape(-1, 2)
    ev = ev[ev[:, 1] != 0]
    if t0 is None: t0 = ev[0, 1]
    rel = [(int(a), int(b - t0)) for a, b in ev]
    if role == 0: print("producer first issues", rel[:8])
    if role < 3:
        # pace inside the first and second tile
        tiles, cur = [], []
        for a, b in rel:
            if a == 1000: continue
            if cur and a < cur[-1][0]: tiles.append(cur); cur = []
            cur.append((a, b))
        tiles.append(cur)
        for ti, t in enumerate(tiles[:3]):
            if len(t) > 8:
                d = np.diff([b for _, b in t[6:]])
                print(f"{name} tile {ti}: start {t[0][1]} end {t[-1][1]} blocks {len(t)} pace median {np.median(d):.0f} mean {d.mean():.0f}")
                if ti == 1 and role == 1: print("   mma tile 1 deltas", np.diff([b for _, b in t]).tolist())
                if ti == 1 and role == 0: print("   producer tile 1 deltas", np.diff([b for _, b in t]).tolist())
    else:
        print(name, rel[:40])
        tiles, cur = [], None
        for a, b in rel:
            if a == 0:
                cur = {0: b}; tiles.append(cur)
            elif cur is not None:
                cur[a] = b
        for t in tiles[1:6]:
            keys = sorted(t)
            print('  tile:', ' '.join(f"{k}:+{t[k] - t[0]}" for k in keys))
        full = [t[0] for t in tiles]
        print('  tile period', np.diff(full)[:10])
s.close()
