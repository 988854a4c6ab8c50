"""GPU check of the tcgen05 GEMM in all operand forms + timing at LLaVA-7B shapes.
Run on the GPU box: python tools/gpu_check_gemm.py > gpurun_out/gemm_check.log
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlaifv_b200 import ops

torch.manual_seed(0)
dev = "cuda"
fails = 0

def check(name, got, ref, tol=2e-2):
    global fails
    got = got.float(); ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    bad = not (err <= tol * scale) or not torch.isfinite(got).all().item()
    print(f"{'FAIL' if bad else 'ok  '} {name}: max_abs_err={err:.4g} ref_max={scale:.4g}", flush=True)
    fails += int(bad)

def run_case(M, N, K, a_mn, b_mn, tile_n=0, bias=False, residual=False, act=0, accumulate=False):
    A = (torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)).bfloat16()
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).bfloat16()
    Af = (A.float().t() if a_mn else A.float())
    Bf = (B.float().t() if b_mn else B.float())
    ref = Af @ Bf.t()
    bs = torch.randn(N, device=dev).bfloat16() if bias else None
    rs = torch.randn(M, N, device=dev).bfloat16() if residual else None
    if bias: ref = ref + bs.float()
    if act == 1: ref = torch.nn.functional.gelu(ref.bfloat16().float())
    if act == 2:
        r = ref.bfloat16().float(); ref = r * torch.sigmoid(1.702 * r)
    if residual: ref = ref.bfloat16().float() + rs.float()
    out = None
    if accumulate:
        out = torch.randn(M, N, device=dev).bfloat16()
        ref = ref + out.float()
    got = ops.gemm(A, B, out, a_mn=a_mn, b_mn=b_mn, bias=bs, residual=rs, act=act, accumulate=accumulate, tile_n=tile_n)
    torch.cuda.synchronize()
    check(f"gemm M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bn={tile_n} bias={bias} res={residual} act={act} acc={accumulate}",
          got, ref, tol=1e-2 if not (act or residual or accumulate) else 2e-2)

try:
    # smallest single tile first
    run_case(128, 128, 64, False, False, 128)
    run_case(128, 256, 64, False, False, 256)
    run_case(128, 128, 256, False, False, 128)
    run_case(256, 512, 512, False, False, 256)
    run_case(128, 128, 64, False, True, 128)
    run_case(128, 128, 64, True, True, 128)
    run_case(256, 512, 512, False, True, 256)
    run_case(256, 512, 512, True, True, 256)
    # ragged edges
    run_case(300, 320, 200, False, False, 0)
    run_case(300, 320, 200, False, True, 0)
    run_case(304, 320, 200, True, True, 0)
    run_case(1135, 4096, 1024, False, False, 0, bias=True, act=1)
    run_case(1000, 1024, 4096, False, False, 0, bias=True, act=2)
    run_case(777, 768, 512, False, False, 0, residual=True)
    run_case(512, 768, 1135, True, True, 0, accumulate=True)
    # 2-CTA kernel (tile_n=512)
    run_case(256, 256, 64, False, False, 512)
    run_case(256, 256, 512, False, False, 512)
    run_case(256, 256, 64, False, True, 512)
    run_case(256, 256, 64, True, True, 512)
    run_case(1000, 768, 512, False, False, 512, bias=True, act=1)
    run_case(777, 768, 512, False, True, 512, residual=True)
    run_case(512, 768, 1135, True, True, 512, accumulate=True)
    run_case(4096, 4096, 1024, False, False, 512)
    run_case(4096, 4096, 1024, False, True, 512)
    run_case(4096, 4096, 1135, True, True, 512)
    run_case(18160, 4096, 512, False, False, 512)
    # many tiles / persistent loop with several tiles per CTA
    run_case(4096, 4096, 1024, False, False, 256)
    run_case(4096, 4096, 1024, False, True, 256)
    run_case(4096, 4096, 1135, True, True, 256)
except Exception as e:
    print("EXCEPTION", repr(e), flush=True)
    fails += 1

def bench(M, N, K, a_mn, b_mn, tile_n=0, iters=10):
    A = (torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)).bfloat16()
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(A, B, out, a_mn=a_mn, b_mn=b_mn, tile_n=tile_n)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm(A, B, out, a_mn=a_mn, b_mn=b_mn, tile_n=tile_n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS comparison (reference number only)
    Af = (A.t().contiguous() if a_mn else A); Bf = (B.t().contiguous() if b_mn else B)
    for _ in range(3): torch.matmul(Af, Bf.t())
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): torch.matmul(Af, Bf.t())
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"perf M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bn={tile_n}: {ms:.3f} ms {tf:.0f} TFLOP/s | cuBLAS {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.0f} TFLOP/s", flush=True)

if fails == 0:
    try:
        Mtok = 18160
        for bn in (256, 512):
            bench(Mtok, 12288, 4096, False, False, bn)
            bench(Mtok, 4096, 11008, False, False, bn)
            bench(Mtok, 11008, 4096, False, True, bn)
            bench(4096, 11008, Mtok, True, True, bn)
            bench(22016, 4096, Mtok, True, True, bn)
        bench(Mtok, 12288, 4096, False, False)      # qkv fwd
        bench(Mtok, 4096, 4096, False, False)       # o fwd
        bench(Mtok, 22016, 4096, False, False)      # gate|up fwd
        bench(Mtok, 4096, 11008, False, False)      # down fwd
        bench(Mtok, 4096, 12288, False, True)       # qkv dgrad
        bench(Mtok, 11008, 4096, False, True)       # down dgrad
        bench(12288, 4096, Mtok, True, True)        # qkv wgrad
        bench(4096, 11008, Mtok, True, True)        # down wgrad
        bench(22016, 4096, Mtok, True, True)        # gate|up wgrad
        bench(Mtok, 32000, 4096, False, False)      # lm_head
        bench(9232, 1024, 1024, False, False)       # clip proj
        bench(9232, 4096, 1024, False, False, 128)
        bench(9232, 4096, 1024, False, False, 256)
        bench(8192, 8192, 8192, False, False)
    except Exception as e:
        print("EXCEPTION", repr(e), flush=True)
        fails += 1
print("FAILS", fails)
sys.exit(1 if fails else 0)
