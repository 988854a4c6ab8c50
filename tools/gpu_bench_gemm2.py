import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlaifv_b200 import ops
dev="cuda"
def bench(M,N,K,a_mn,b_mn,tile_n,iters=10):
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
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bn={tile_n}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
Mtok=18160
for bn in (256,512):
    bench(Mtok,12288,4096,False,False,bn)
    bench(Mtok,4096,4096,False,False,bn)
    bench(Mtok,22016,4096,False,False,bn)
    bench(Mtok,4096,11008,False,False,bn)
    bench(Mtok,11008,4096,False,True,bn)
    bench(4096,11008,Mtok,True,True,bn)
    bench(22016,4096,Mtok,True,True,bn)
    bench(Mtok,32000,4096,False,False,bn)
    bench(8192,8192,8192,False,False,bn)
