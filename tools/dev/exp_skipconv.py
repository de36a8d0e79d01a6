"""Timing experiment (BAE_TC_DEBUG build): step time at Madelon, 64 chains, with the hi/lo conversion of A / B / both skipped
(BAE_TC_DBGFLAGS 2 / 4 / 1; results are wrong, timing only) - the upper bound of what pre-converted operand images can buy."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
from bench import SHAPES, synth
dims4, ntr, nte = SHAPES["madelon"]
train, test = synth(dims4, ntr, nte)
C = 64
s = DeviceSampler(dims4, train, test, C, 8, 100, np.tile(2.0 ** (np.arange(8) / 7), C // 8))
rng = np.random.default_rng(0)
for j in range(C): s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
s.run(5); s.synchronize()
ms = []
for _ in range(4):
    s.run(5); s.synchronize(); ms.append(s.last_ms())
print("flags", os.environ.get("BAE_TC_DBGFLAGS", "0"), "ms per round", [round(x, 2) for x in ms])
s.close()
