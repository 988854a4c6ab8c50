"""ncu launch list (csv with gpu__time_duration.sum per launch) -> per-kernel shares of the step.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launch_shares.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3,
              "second": 1e3}[unit]
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").strip()
    rows.append((name, ms))
tot = sum(ms for _, ms in rows)
agg = defaultdict(lambda: [0.0, 0])
for n, ms in rows:
    agg[n][0] += ms
    agg[n][1] += 1
print(f"launches {len(rows)} total {tot:.1f} ms (ncu: cold-cache, serialised — compare shares)")
for n, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * ms / tot:6.2f}%  {ms:9.2f} ms  n={c:5d}  {n}")
gemm = sum(ms for n, (ms, c) in agg.items() if "gemm" in n)
print(f"all GEMM kernels: {100 * gemm / tot:.2f}% of the step")
