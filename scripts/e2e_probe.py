"""Host-API latency probe: python scripts/e2e_probe.py -- per-call latency of cvb_akaze_extract_batch / cvb_match_symmetric with
1, 4 and 12 host threads (one context each), pinned host buffers; prints frames/s and the mean call times."""
import ctypes as C, os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import cv_b200
from cv_b200._lib import KP_DTYPE
from tests.synth import synth_frame, warp_frame
W,H=1920,1080
a=synth_frame(0); pair=np.stack([a,warp_frame(a,1000)])
cap=8192
cfg=cv_b200.AkazeConfig(maximum_features=5000).to_c()
def mk():
    cx=cv_b200.Context(0)
    o=dict(img=torch.from_numpy(pair).pin_memory(), kp=torch.empty(2*cap*KP_DTYPE.itemsize,dtype=torch.uint8).pin_memory(), desc=torch.empty(2*cap*64,dtype=torch.uint8).pin_memory(), n=torch.zeros(2,dtype=torch.int32).pin_memory(), pairs=torch.empty(cap*2,dtype=torch.int32).pin_memory(), np_=C.c_uint32())
    return cx,o
def step(cx,o,t):
    t0=time.perf_counter()
    cx.check(cx.lib.cvb_akaze_extract_batch(cx.handle,C.byref(cfg),o['img'].data_ptr(),2,W,H,o['kp'].data_ptr(),o['desc'].data_ptr(),cap,o['n'].data_ptr()))
    t1=time.perf_counter()
    na,nb=int(o['n'][0]),int(o['n'][1])
    t2=time.perf_counter()
    cx.check(cx.lib.cvb_match_symmetric(cx.handle,o['desc'].data_ptr(),na,o['desc'].data_ptr()+cap*64,nb,24,o['pairs'].data_ptr(),cap,C.byref(o['np_'])))
    t3=time.perf_counter()
    t[0]+=t1-t0; t[1]+=t2-t1; t[2]+=t3-t2
for NT in (1,4,12):
    cs=[mk() for _ in range(NT)]
    for cx,o in cs:
        for _ in range(3): step(cx,o,[0,0,0])
    K=20
    ts=[[0,0,0] for _ in range(NT)]
    def work(i):
        for _ in range(K): step(cs[i][0],cs[i][1],ts[i])
    th=[threading.Thread(target=work,args=(i,)) for i in range(NT)]
    t0=time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt=time.perf_counter()-t0
    tt=np.array(ts).sum(0)/(NT*K)*1e3
    print(f"threads {NT}: {2*NT*K/dt:.0f} fps; per step ms: extract {tt[0]:.3f} python {tt[1]:.3f} match {tt[2]:.3f}")
