import os, sys, numpy as np
os.environ["BAE_TC_TRACE"] = "1"; os.environ["BAE_PHASE_LAUNCH"] = "1"
sys.path.insert(0, '.')
from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
from bayesianautoencoders_b200 import _native as N
from bench import SHAPES, synth
dims4, ntr, nte = SHAPES["madelon"]
train, test = synth(dims4, ntr, nte)
C = 16
s = DeviceSampler(dims4, train, test, C, 8, 100, np.tile(2.0 ** (np.arange(8) / 7), C // 8))
rng = np.random.default_rng(0)
for j in range(C): s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
buf = np.zeros(4096, np.int64)
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 1))
s.step(1); s.synchronize()
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 0))
t0 = None
ev = buf[2 * 1024:3 * 1024].reshape(-1, 2)
for g in range(3):
    e = ev[g * 170:(g + 1) * 170]; e = e[e[:, 1] != 0]
    print("conv group", g, [(int(a), int(b - buf[1])) for a, b in e[:40]])
for role, name in enumerate(["producer", "mma"]):
    ev = buf[role * 1024:(role + 1) * 1024].reshape(-1, 2)
    ev = ev[ev[:, 1] != 0]
    if t0 is None: t0 = ev[0, 1]
    print(name, [(int(a), int(b - t0)) for a, b in ev[:64]])
s.close()
