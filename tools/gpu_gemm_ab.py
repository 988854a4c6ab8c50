import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlaifv_b200 import ops, lib
dev="cuda"; L=lib.load()
Mtok=18160
shapes=[(Mtok,12288,4096,False,False),(Mtok,4096,11008,False,False),(Mtok,11008,4096,False,True),(22016,4096,Mtok,True,True),(Mtok,22016,4096,False,False)]
bufs=[]
for (M,N,K,a_mn,b_mn) in shapes:
    A=(torch.randn(K,M,device=dev) if a_mn else torch.randn(M,K,device=dev)).bfloat16()
    B=(torch.randn(K,N,device=dev) if b_mn else torch.randn(N,K,device=dev)).bfloat16()
    bufs.append((A,B,torch.empty(M,N,device=dev,dtype=torch.bfloat16)))
def run_all(iters=6):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for (M,N,K,a_mn,b_mn),(A,B,C) in zip(shapes,bufs):
            ops.gemm(A,B,C,a_mn=a_mn,b_mn=b_mn,tile_n=256)
    e1.record(); torch.cuda.synchronize()
    fl=sum(2.0*M*N*K for (M,N,K,_,_) in shapes)*iters
    ms=e0.elapsed_time(e1)
    return fl/ms/1e9
for mode in (0,4): 
    L.rlaifv_gemm_set_tuning(16,mode); run_all(3)
res={0:[],4:[]}
for rep in range(6):
    for mode in (0,4):
        L.rlaifv_gemm_set_tuning(16,mode)
        res[mode].append(run_all())
L.rlaifv_gemm_set_tuning(16,0)
print("dynamic TFLOP/s:", [round(x) for x in res[0]], "mean", round(sum(res[0])/6))
print("static  TFLOP/s:", [round(x) for x in res[4]], "mean", round(sum(res[4])/6))
