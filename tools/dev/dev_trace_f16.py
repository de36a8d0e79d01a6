"""Timeline of one GEMM phase of CTA 0 on the 3xFP16 engine (BAE_TC_TRACE=1, BAE_TC_TRACE_PHASE=<phase index>): K-block pace of
the TMA producer and the MMA issuer, and per converter group the wait and conversion time of every K block it takes."""
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
print("phase", ph, "step ms", s.last_ms())
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 0))
def events(role, lo=0, hi=512):
    ev = buf[role * 1024:(role + 1) * 1024].reshape(-1, 2)[lo:hi]
    return [(int(a), int(b)) for a, b in ev if b != 0]
p = events(0); m = [e for e in events(1) if e[0] != 1000]
t0 = p[0][1]
for name, ev in (("producer", p), ("mma", m)):
    t = np.array([b for _, b in ev]) - t0
    d = np.diff(t)
    print(f"{name}: {len(t)} K blocks, first {t[0]}, last {t[-1]}, pace median {np.median(d):.0f} mean {d.mean():.0f}; first 40 deltas {d[:40].tolist()}")
for g, (lo, hi) in enumerate(((0, 170), (170, 340), (340, 510))):
    ev = events(2, lo, hi)
    w_lo, w_full, conv, idle, t7, tend = [], [], [], [], None, None
    t8 = t0k = tf = 0   # (the 3xFP16 converter loop stamps only block start / end)
    fence, arr = [], []
    for a, b in ev:     # per K block: 7000 + kb (arrived), 8000 + kb (lo slot free), kb (TMA bytes landed), 2000 + kb (converted)
        if 7000 <= a < 8000:
            t7 = b
            if tend is not None: idle.append(b - tend)
        elif 8000 <= a < 9000: t8 = b; w_lo.append(b - t7)
        elif a < 1000: t0k = b; w_full.append(b - t8 if t8 else 0)
        elif 2000 <= a < 3000: conv.append(b - t0k); tend = b
        elif 3000 <= a < 4000: fence.append(b - tend); tf = b
        elif 4000 <= a < 5000: arr.append(b - tf); tend = b
    if conv:
        print(f"converter group {g}: {len(conv)} blocks; lo-slot wait median {np.median(w_lo):.0f} mean {np.mean(w_lo):.0f}; TMA wait median "
              f"{np.median(w_full):.0f} mean {np.mean(w_full):.0f}; conversion median {np.median(conv):.0f} mean {np.mean(conv):.0f}; "
              f"proxy fence median {np.median(fence):.0f}; arrivals median {np.median(arr):.0f}; skip two blocks median {np.median(idle):.0f}")
        print("   first 12 (lo wait, TMA wait, conversion):", list(zip(w_lo[:12], w_full[:12], conv[:12])))
e = events(3)
t = [b - t0 for a, b in e if a == 0]
print("epilogue tile starts", t[:8], "period", np.diff(t)[:8].tolist())
tiles, cur = [], None
for a, b in e:
    if a == 0:
        cur = {0: b}; tiles.append(cur)
    elif cur is not None and a not in cur:
        cur[a] = b
for tl in tiles[1:5]:
    print("  epilogue tile:", " ".join(f"{k}:+{tl[k] - tl[0]}" for k in sorted(tl)))
s.close()
