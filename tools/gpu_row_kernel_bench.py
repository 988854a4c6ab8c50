"""Achieved HBM bandwidth of the row kernels at the config-(b) shapes, timed with CUDA events on the launching stream
(20 launches after 3 warm-ups; every operand is far larger than the 126 MB L2). Algorithmic bytes as in
tools/row_kernel_bandwidth.py; peak = MEASURED_PEAKS.json hbm_gbs.

    python tools/gpu_row_kernel_bench.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlaifv_b200 import ops

dev = "cuda"
M, H, F, V, T, NS, NH = 18160, 4096, 11008, 32000, 1135, 16, 32
BF, F32 = torch.bfloat16, torch.float32
peak = 6572.2
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("hbm_gbs", peak)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


x = torch.randn(M, H, device=dev).to(BF)
dy = torch.randn(M, H, device=dev).to(BF)
dres = torch.randn(M, H, device=dev).to(BF)
w = torch.ones(H, device=dev, dtype=BF)
y = torch.empty_like(x)
dx = torch.empty_like(x)
rstd = torch.empty(M, device=dev, dtype=F32)
dw = torch.zeros(H, device=dev, dtype=BF)
gu = torch.randn(M, 2 * F, device=dev).to(BF)
act = torch.empty(M, F, device=dev, dtype=BF)
dact = torch.randn(M, F, device=dev).to(BF)
dgu = torch.empty_like(gu)
qkv = torch.randn(M, 3 * H, device=dev).to(BF)
cos = torch.randn(T, 128, device=dev).to(BF)
sin = torch.randn(T, 128, device=dev).to(BF)
o = torch.randn(M, H, device=dev).to(BF)
delta = torch.empty(NS, NH, T, device=dev, dtype=F32)
lse = torch.zeros(NS, NH, T, device=dev, dtype=F32)
ops.rmsnorm_fwd(x, w, 1e-5, out=y, rstd=rstd)
rows = [
    ("rmsnorm_fwd", lambda: ops.rmsnorm_fwd(x, w, 1e-5, out=y, rstd=rstd), 2 * M * H * 2),
    ("rmsnorm_bwd (+residual grad)", lambda: ops.rmsnorm_bwd(dy, x, w, rstd, dx, dw, dres=dres, dw_accumulate=False), 4 * M * H * 2),
    ("swiglu_fwd", lambda: ops.swiglu_fwd(gu, act), 3 * M * F * 2),
    ("swiglu_bwd", lambda: ops.swiglu_bwd(gu, dact, dgu), 5 * M * F * 2),
    ("rope_fwd", lambda: ops.rope_fwd(qkv, cos, sin, T, NH, 128), 2 * M * 2 * H * 2),
    ("rope_bwd (bf16 dq in place)", lambda: ops.rope_bwd(qkv, None, cos, sin, T, NH, 128), 2 * M * 2 * H * 2),
]
print("%-34s %10s %10s %8s" % ("kernel", "ms/launch", "GB/s", "of %.0f" % peak))
for name, fn, nbytes in rows:
    ms = timeit(fn)
    print("%-34s %10.4f %10.0f %7.1f%%" % (name, ms, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / peak))
# attention delta (launched inside the split backward: time the backward's first kernel through the profiler-free
# route: a tiny sequence would not be HBM-bound, so call the C entry indirectly via the split backward and subtract?)
# -> measured separately by ncu (profiles/*launch*): listed there.
