"""Achieved (algorithmic) HBM bandwidth of the HBM-bound kernels from the one-step ncu launch list:
bytes the kernel must move at the BASELINE config-(b) shapes (M = 18 160 tokens, H = 4096, F = 11008, V = 32000)
divided by its `gpu__time_duration.sum`. usage: python tools/row_kernel_bandwidth.py profiles/r01c_launches_one_step.csv"""
import csv
import re
import sys
from collections import defaultdict

M, H, F, V = 18160, 4096, 11008, 32000
PEAK = 6572.2   # measured copy bandwidth of this pool's B200s (MEASURED_PEAKS.json / B200_PROFILING.md), GB/s
MB = 1e6
# kernel -> (algorithmic bytes per launch, what is counted); only launches of the dominant (decoder-sized) grid are averaged
ALG = {
    "rmsnorm_fwd_kernel": (2 * M * H * 2, "read x + write y"),
    "rmsnorm_bwd_kernel": (4 * M * H * 2, "read dy, x, residual-grad + write dx"),
    "swiglu_fwd_kernel": (3 * M * F * 2, "read gate|up (2F) + write act (F)"),
    "swiglu_bwd_kernel": (5 * M * F * 2, "read gate|up (2F), d_act (F) + write d gate|up (2F)"),
    "rope_fwd_kernel": (2 * M * 2 * H * 2, "q and k blocks of the fused qkv buffer, read + write in place"),
    "rope_bwd_kernel": (M * H * 4 + M * H * 2 + 2 * M * H * 2, "read fp32 dQ, write bf16 dq, dk read + write in place"),
    "attention_delta_kernel": (2 * M * H * 2, "read O and dO"),
    "logp_fwd_kernel": (M * V * 2, "read the bf16 logits once"),
    "logp_bwd_kernel": (int(M * V * 2 * (1 + 512 / 1135)), "write every row; read only the 512 of 1135 supervised rows"),
    "adamw_kernel<0>": (28 * 202383360, "fp32 master/m/v read + write, bf16 grad read, bf16 param write (layer bucket)"),
}
rows = defaultdict(list)
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    ms = v * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}[r["Metric Unit"]]
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").replace("b200::", "").strip()
    rows[name].append(ms)
print(f"{'kernel':26s} {'launches':>8s} {'ms/launch':>10s} {'alg MB':>9s} {'GB/s':>8s} {'of %.0f' % PEAK:>8s}  counted bytes")
for k, (b, what) in ALG.items():
    ts = rows.get(k, [])
    if not ts:
        continue
    big = [t for t in ts if t > 0.5 * max(ts)]          # drop the small launches (CLIP-sized / tail buckets)
    ms = sum(big) / len(big)
    gbs = b / (ms * 1e-3) / 1e9
    print(f"{k:26s} {len(big):8d} {ms:10.4f} {b / MB:9.0f} {gbs:8.0f} {100 * gbs / PEAK:7.1f}%  {what}")
print("(durations: ncu gpu__time_duration per launch, cold L2, serialised; the same kernels inside the overlapped step "
      "run beside GEMMs)")
