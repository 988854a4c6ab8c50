import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlaifv_b200 import ops, lib
dev="cuda"
Mtok=18160
shapes={"fwd_down K=11008":(Mtok,4096,11008,False,False),"dgrad_qkv K=12288":(Mtok,4096,12288,False,True),
        "dgrad_gu K=22016":(Mtok,4096,22016,False,True),"lm_head fwd N=32000":(Mtok,32000,4096,False,False),
        "lm_head dgrad K=32000":(Mtok,4096,32000,False,True),"fwd_qkv":(Mtok,12288,4096,False,False),"dgrad_down":(Mtok,11008,4096,False,True),
        "wgrad_qkv":(12288,4096,Mtok,True,True),"wgrad_gu":(22016,4096,Mtok,True,True),"wgrad_down":(4096,11008,Mtok,True,True),
        "fwd_o":(Mtok,4096,4096,False,False)}
bufs={}
for k,(M,N,K,a_mn,b_mn) in shapes.items():
    A=(torch.randn(K,M,device=dev) if a_mn else torch.randn(M,K,device=dev)).bfloat16()
    B=(torch.randn(K,N,device=dev) if b_mn else torch.randn(N,K,device=dev)).bfloat16()
    bufs[k]=(A,B,torch.empty(M,N,device=dev,dtype=torch.bfloat16))
def run(k,bn,iters=5):
    M,N,K,a_mn,b_mn=shapes[k]; A,B,C=bufs[k]
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm(A,B,C,a_mn=a_mn,b_mn=b_mn,tile_n=bn)
    e1.record(); torch.cuda.synchronize()
    return 2.0*M*N*K*iters/e0.elapsed_time(e1)/1e9
res={k:{256:[],512:[]} for k in shapes}
for k in shapes:
    for bn in (256,512): run(k,bn,2)
for rep in range(5):
    for k in shapes:
        for bn in (256,512): res[k][bn].append(run(k,bn))
for k in shapes:
    a=sum(res[k][256])/5; b=sum(res[k][512])/5
    print(f"{k:24s} 1-CTA {a:5.0f}  2-CTA {b:5.0f}  ratio {b/a:.3f}")
