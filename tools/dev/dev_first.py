import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import bae_oracle as bo
from bayesianautoencoders_b200 import DeviceSampler
from tests.parity_util import synth_data, rel
for name,(dims4,ntr,nte) in {"mid":((24,20,16,12),96,40),"swiss":((3,10,5,2),3750,1250),"madelon":((500,450,400,300),2000,1800)}.items():
    train,test=synth_data(dims4,ntr,nte)
    s=DeviceSampler(dims4,train,test,1,1,20,[1.0])
    L=bo.Layout(dims4)
    w=(np.random.default_rng(5).standard_normal(L.w_size)).astype(np.float32); w[L.nbt_index]=3
    t=time.time(); r=s.eval(0,0,w=w,grads=True,out=True); print(name,'eval time',time.time()-t)
    th=w.astype(np.float64)
    out,cache=bo.forward(L,th,train.astype(np.float64),update_buffers=False)
    g=bo.backward(L,th,cache)
    sse=float(np.sum((out-train)**2))
    print(name,'sse',r['sse'],sse,(r['sse']-sse)/sse,'out rel',rel(r['out'],out))
    for k in bo.SORTED_KEYS:
        if k in bo.BUFFER_KEYS: continue
        sl=L.slice(k); print('   ',k,rel(r['grads'][sl],g[sl]), float(np.abs(g[sl]).max()))
    s.close()
