"""Phase durations INSIDE the persistent step kernel (debug, BAE_TC_TRACE=1): globaltimer stamps of CTA 0 after each
grid barrier of the 2nd iteration of one launch."""
import os, sys, numpy as np
os.environ["BAE_TC_TRACE"] = "1"
sys.path.insert(0, '.')
from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
from bayesianautoencoders_b200 import _native as N
from bench import SHAPES, synth
dims4, ntr, nte = SHAPES["madelon"]
train, test = synth(dims4, ntr, nte)
C = 64
s = DeviceSampler(dims4, train, test, C, 8, 100, np.tile(2.0 ** (np.arange(8) / 7), C // 8))
rng = np.random.default_rng(0)
for j in range(C): s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
buf = np.zeros(4096, np.int64)
s.run(5); s.synchronize()
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 1))
s.run(5); s.synchronize()
print("launch ms", s.last_ms())
N.check(s.lib.bae_debug_trace(s._h, buf.ctypes.data, 0))
t = buf[3584:3584 + 42]
names = ['PRO','F1','F2','F3','BNF','F4','F5','F6','B6','B5','B4','BNB','B3','B2','B1','ADAM1','F1b','F2b','F3b','BNFb','F4b','F5b','F6b','B6b','B5b','B4b','BNBb','B3b','B2b','B1b','ADAM2','T1','T2','T3','BNT','T4','T5','T6','EPI','SW1','SW2','SW3']
d = np.diff(t[:42]) / 1000.0
print({n: round(float(x)) for n, x in zip(names[1:], d)})
print("sum us", float(d[:38].sum()))
s.close()
