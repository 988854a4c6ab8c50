"""L2 behaviour of the CTA-pair GEMM on the decoder's dgrad / wgrad shapes: raster group size, m- vs n-grouped walk,
static vs dynamic tile draw, and TMA L2 eviction hints (rlaifv_gemm_set_tuning / rlaifv_gemm_set_l2).

  timing (CUDA events, in-process A/B, each config timed back to back for >= 0.3 s so the power cap settles):
      python tools/gpu_gemm_l2_sweep.py
  DRAM bytes (one launch per config, in the order printed by --list):
      ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_elapsed.max \
          -k regex:gemm2 --csv --log-file gpurun_out/l2_sweep.csv python tools/gpu_gemm_l2_sweep.py --once
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlaifv_b200 import lib, ops

L = lib.load()
dev = "cuda"
M_TOK = 16 * 1135
SHAPES = {   # name: (M, N, K, a_mn, b_mn)   C[M,N] = op(A) op(B)^T
    "dgrad_qkv": (M_TOK, 4096, 12288, False, True),
    "wgrad_qkv": (12288, 4096, M_TOK, True, True),
    "dgrad_gu": (M_TOK, 4096, 22016, False, True),
    "wgrad_gu": (22016, 4096, M_TOK, True, True),
    "wgrad_down": (4096, 11008, M_TOK, True, True),
    "fwd_qkv": (M_TOK, 12288, 4096, False, False),
}
# (label, group_m, debug bits (4 = static schedule), l2 bits)
CONFIGS = [
    ("base g16", 16, 0, 0),
    ("auto (shipped default)", 16, 0, -1),
    ("g8", 8, 0, 0),
    ("g4", 4, 0, 0),
    ("g32", 32, 0, 0),
    ("g8 static", 8, 4, 0),
    ("g16 static", 16, 4, 0),
    ("ng8", 8, 0, 0x40),
    ("ng4", 4, 0, 0x40),
    ("g8 A=last B=first", 8, 0, 2 | (1 << 2)),
    ("g8 A=last", 8, 0, 2),
    ("g16 A=last B=first", 16, 0, 2 | (1 << 2)),
    ("g8 B=last A=first", 8, 0, 1 | (2 << 2)),
    ("ng8 B=last A=first", 8, 0, 0x40 | 1 | (2 << 2)),
    ("g8 C=first", 8, 0, 1 << 4),
    ("g8 A=last B=first C=first", 8, 0, 2 | (1 << 2) | (1 << 4)),
]


def operands(M, N, K, a_mn, b_mn):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16() * 0.05
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16() * 0.05
    c = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    return a, b, c


def main():
    once = "--once" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    names = only or list(SHAPES)
    if "--list" in sys.argv:
        i = 0
        for n in names:
            for cfg in CONFIGS:
                print(i, n, cfg[0])
                i += 1
        return
    for n in names:
        M, N, K, a_mn, b_mn = SHAPES[n]
        a, b, c = operands(M, N, K, a_mn, b_mn)
        fl = 2.0 * M * N * K
        ref = None
        for label, gm, dbg, l2 in CONFIGS:
            L.rlaifv_gemm_set_tuning(gm, dbg)
            L.rlaifv_gemm_set_l2(l2)
            run = lambda: ops.gemm(a, b, out=c, a_mn=a_mn, b_mn=b_mn)
            if once:
                run()
                torch.cuda.synchronize()
                continue
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            if ref is None:
                ref = c.clone()
            else:
                assert torch.equal(ref, c), (n, label)          # raster / hints must not change a single bit
            reps = max(10, int(0.3 / (fl / 1.4e15)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = []
            for _ in range(2):
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                best.append(e0.elapsed_time(e1) / reps)
            ms = min(best)
            print("%-10s %-28s %.4f ms  %6.0f TFLOP/s" % (n, label, ms, fl / ms / 1e9), flush=True)
        L.rlaifv_gemm_set_tuning(16, 0)
        L.rlaifv_gemm_set_l2(-1)
        del a, b, c


if __name__ == "__main__":
    main()
