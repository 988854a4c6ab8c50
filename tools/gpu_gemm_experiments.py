import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlaifv_b200 import ops, lib
dev="cuda"
def bench(M,N,K,a_mn,b_mn,tile_n,tag,iters=10):
    A=(torch.randn(K,M,device=dev) if a_mn else torch.randn(M,K,device=dev)).bfloat16()
    B=(torch.randn(K,N,device=dev) if b_mn else torch.randn(N,K,device=dev)).bfloat16()
    out=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(A,B,out,a_mn=a_mn,b_mn=b_mn,tile_n=tile_n)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm(A,B,out,a_mn=a_mn,b_mn=b_mn,tile_n=tile_n)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    print(f"{tag:28s} M={M} N={N} K={K} mn=({int(a_mn)},{int(b_mn)}) bn={tile_n}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
Mtok=18160
shapes=[(Mtok,12288,4096,False,False),(Mtok,4096,11008,False,False),(Mtok,11008,4096,False,True),(22016,4096,Mtok,True,True)]
L=lib.load()
for gm in (8,16,32,64,142):
    L.rlaifv_gemm_set_tuning(gm,0)
    for sh in shapes: bench(*sh,256,f"group_m={gm}")
L.rlaifv_gemm_set_tuning(16,0)
for dbg,name in ((1,"no-store"),(3,"no-tmem-ld/no-store")):
    L.rlaifv_gemm_set_tuning(16,dbg)
    for sh in shapes:
        bench(*sh,256,name); bench(*sh,512,name+" 2cta")
L.rlaifv_gemm_set_tuning(16,0)
