import os, sys, torch, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lina_speech_b200 import _lib as L
dev="cuda"; B,n,d=32,128,1024
q=torch.randn(B,d,device=dev,dtype=torch.bfloat16); k=torch.randn(B,n,d,device=dev,dtype=torch.bfloat16); pe=torch.randn(n,d,device=dev,dtype=torch.bfloat16)
lw=torch.ones(d,device=dev,dtype=torch.bfloat16); lb=torch.zeros(d,device=dev,dtype=torch.bfloat16)
att=torch.empty(B,2,1,n,device=dev,dtype=torch.bfloat16); x=torch.empty(B,d,device=dev,dtype=torch.bfloat16)
lib=L.lib()
def f():
    return lib.lina_cross_att_step(L.ptr(q), d, L.ptr(lw), L.ptr(lb), 1e-5, L.ptr(k), n*d, L.ptr(pe), 0, L.ptr(att), 2*n, L.ptr(x), d, B, n, d, 1/math.sqrt(d), L.dt(q), L.stream(q))
def g():
    w = (q.unsqueeze(1) @ k.transpose(-2,-1)) * (1/math.sqrt(d)); w=torch.softmax(w,-1); return w @ pe
for fn,name in ((f,'fused kernel'),(g,'torch 4 ops')):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1)/200*1000, 'us per call (back-to-back launches)')
gr=torch.cuda.CUDAGraph()
for fn,name in ((f,'fused kernel'),(g,'torch 4 ops')):
    s=torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); 
    torch.cuda.synchronize()
    gr=torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20): fn()
    gr.replay(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(name, 'in graph:', e0.elapsed_time(e1)/400*1000, 'us per call')
