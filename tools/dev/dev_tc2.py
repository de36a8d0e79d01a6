import sys, numpy as np
sys.path.insert(0, '.')
from oracle import bae_oracle as bo
from bayesianautoencoders_b200 import DeviceSampler
from tests.parity_util import synth_data, rel
dims4, ntr, nte = ((256, 256, 256, 256), 512, 256)
train, test = synth_data(dims4, ntr, nte)
L = bo.Layout(dims4)
w = (np.random.default_rng(5).standard_normal(L.w_size) * 0.3).astype(np.float32); w[L.nbt_index] = 3
th = w.astype(np.float64)
out, c = bo.forward(L, th, train.astype(np.float64), update_buffers=False)
N, d0 = train.shape
dout = 2.0 * (out - train) / (N * d0)
W6 = L.view(th, "decode.5.weight")
d5 = (dout @ W6) * c["h5"] * (1 - c["h5"])
s = DeviceSampler(dims4, train, test, 1, 1, 20, [1.0], gemm_path=2)
r = s.eval(0, 0, w=w, grads=True)
E = 1
def get(name, d):
    hi = s.debug_buffer(name, E, N); lo = s.debug_buffer(name, E, N, lo_part=True)
    return (hi.astype(np.float64) + lo)[:, :d], hi, lo
H5, hi, lo = get("H5", 256); print("H5 rel", rel(H5, c["h5"]), "ones col", hi[:3, 256], lo[:3, 256])
DO, hi, lo = get("DOUT", 256); print("DOUT rel", rel(DO, dout))
D5, hi, lo = get("D5", 256); print("D5 rel", rel(D5, d5), np.abs(D5).max(), np.abs(d5).max())
print("D5 dev sample", D5[:2, :4], "ref", d5[:2, :4])
g = np.zeros(4096, np.float32)
s.close()
