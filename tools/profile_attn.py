"""Three launches each of the attention forward / backward kernels at the config-(b) layer shape, for

    ncu --set full --clock-control none --import-source on -k regex:attention_ -c 6 -o gpurun_out/attn python tools/profile_attn.py
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlaifv_b200 import lib, ops

lib.load().rlaifv_attention_set_variant(int(os.environ.get("RLAIFV_ATT_VARIANT", "1")))
nseq, S, nh, D = 16, 1135, 32, 128
H = nh * D
torch.manual_seed(0)
qkv = torch.randn(nseq * S, 3 * H, device="cuda").bfloat16()
q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
scale = 1.0 / math.sqrt(D)
out, lse = ops.attention_fwd(q, k, v, nseq, S, nh, D, True, scale)
d_out = torch.randn(nseq * S, H, device="cuda").bfloat16()
dq32 = torch.zeros(nseq * S, H, device="cuda", dtype=torch.float32)
dqkv = torch.zeros(nseq * S, 3 * H, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention_fwd(q, k, v, nseq, S, nh, D, True, scale, out, lse)
if "--fwd-ab" in sys.argv:                  # one launch of each forward variant, for a side-by-side capture
    for var in (0, 1):
        lib.load().rlaifv_attention_set_variant(var)
        ops.attention_fwd(q, k, v, nseq, S, nh, D, True, scale, out, lse)
if "--bwd" in sys.argv:
    for _ in range(2):
        ops.attention_bwd(q, k, v, out, d_out, lse, nseq, S, nh, D, scale, dq32, dqkv[:, H:2 * H], dqkv[:, 2 * H:])
if "--bwd-split" in sys.argv:
    for _ in range(2):
        ops.attention_bwd_split(q, k, v, out, d_out, lse, nseq, S, S, nh, D, True, scale, dqkv[:, :H], dqkv[:, H:2 * H],
                                dqkv[:, 2 * H:])
torch.cuda.synchronize()
print("done")
