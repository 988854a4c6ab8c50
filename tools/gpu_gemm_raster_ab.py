import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlaifv_b200 import ops, lib
dev="cuda"; L=lib.load()
Mtok=18160
shapes={"fwd_qkv":(Mtok,12288,4096,False,False),"fwd_down":(Mtok,4096,11008,False,False),"fwd_gu":(Mtok,22016,4096,False,False),
        "dgrad_qkv":(Mtok,4096,12288,False,True),"dgrad_gu":(Mtok,4096,22016,False,True),"dgrad_down":(Mtok,11008,4096,False,True),
        "wgrad_qkv":(12288,4096,Mtok,True,True),"wgrad_gu":(22016,4096,Mtok,True,True),"wgrad_down":(4096,11008,Mtok,True,True)}
bufs={}
for k,(M,N,K,a_mn,b_mn) in shapes.items():
    A=(torch.randn(K,M,device=dev) if a_mn else torch.randn(M,K,device=dev)).bfloat16()
    B=(torch.randn(K,N,device=dev) if b_mn else torch.randn(N,K,device=dev)).bfloat16()
    bufs[k]=(A,B,torch.empty(M,N,device=dev,dtype=torch.bfloat16))
def run(k,iters=5):
    M,N,K,a_mn,b_mn=shapes[k]; A,B,C=bufs[k]
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm(A,B,C,a_mn=a_mn,b_mn=b_mn,tile_n=256)
    e1.record(); torch.cuda.synchronize()
    return 2.0*M*N*K*iters/e0.elapsed_time(e1)/1e9
groups=[1,2,4,8,16,32]
res={k:{g:[] for g in groups} for k in shapes}
for g in groups:
    L.rlaifv_gemm_set_tuning(g,0)
    for k in shapes: run(k,2)
for rep in range(4):
    for k in shapes:
        for g in groups:
            L.rlaifv_gemm_set_tuning(g,0)
            res[k][g].append(run(k))
L.rlaifv_gemm_set_tuning(16,0)
for k in shapes:
    print(f"{k:11s}", "  ".join(f"g{g}:{sum(res[k][g])/len(res[k][g]):5.0f}" for g in groups))
