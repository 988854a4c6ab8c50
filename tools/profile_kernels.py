"""Short single-GPU command for `ncu --set full`: a few launches of the dominant kernels at the
BASELINE config-(b) shapes (GEMM fwd/dgrad/wgrad, attention fwd/bwd)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlaifv_b200 import ops
dev = "cuda"
M, H, F = 18160, 4096, 11008
x = torch.randn(M, H, device=dev).bfloat16()
w = torch.randn(3 * H, H, device=dev).bfloat16() * 0.02
dy = torch.randn(M, 3 * H, device=dev).bfloat16()
y = torch.empty(M, 3 * H, device=dev, dtype=torch.bfloat16)
dx = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
dw = torch.empty(3 * H, H, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(x, w, y)                                   # forward  (K-major, K-major)
    ops.gemm(dy, w, dx, b_mn=True)                      # dgrad    (K-major, MN-major)
    ops.gemm(dy, x, dw, a_mn=True, b_mn=True)           # wgrad    (MN-major, MN-major)
# the two largest launches of the backward: gate|up dgrad (K = 22016) and the down-projection dgrad with the SwiGLU
# backward in its epilogue
w_gu = torch.randn(2 * F, H, device=dev).bfloat16() * 0.02
w_dn = torch.randn(H, F, device=dev).bfloat16() * 0.02
dgu = torch.randn(M, 2 * F, device=dev).bfloat16()
gu = torch.randn(M, 2 * F, device=dev).bfloat16()
dx3 = torch.randn(M, H, device=dev).bfloat16()
dgu_out = torch.empty(M, 2 * F, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm(dgu, w_gu, dx, b_mn=True)                  # dgrad gate|up
    ops.gemm_swiglu_bwd(dx3, w_dn, gu, dgu_out)         # dgrad down + SwiGLU backward epilogue
del w_gu, w_dn, dgu, gu, dx3, dgu_out
nseq, S, nh, D = 16, 1135, 32, 128
qkv = torch.randn(nseq * S, 3 * H, device=dev).bfloat16()
q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
scale = 1 / math.sqrt(D)
out, lse = ops.attention_fwd(q, k, v, nseq, S, nh, D, True, scale)
do = torch.randn(nseq * S, H, device=dev).bfloat16()
dq32 = torch.zeros(nseq * S, H, device=dev, dtype=torch.float32)
dqkv = torch.zeros(nseq * S, 3 * H, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.attention_fwd(q, k, v, nseq, S, nh, D, True, scale, out, lse)
    ops.attention_bwd_split(q, k, v, out, do, lse, nseq, S, S, nh, D, True, scale, dqkv[:, :H], dqkv[:, H:2 * H],
                            dqkv[:, 2 * H:])
torch.cuda.synchronize()
print("done")
