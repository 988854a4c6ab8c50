"""`ncu --set full` report -> small JSON (per launch: duration, DRAM bytes, tensor-pipe activity, L2 hit rate).
usage: python tools/summarize_ncu_full.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_summary.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__cluster_size"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
rd = list(csv.reader(io.StringIO(raw.stdout)))
hdr, units, body = rd[0], rd[1], rd[2:]
idx = {h: i for i, h in enumerate(hdr)}
out = {"units": {k: units[idx[k]] for k in KEEP if k in idx}, "kernels": []}
for r in body:
    e = {"Kernel Name": r[idx["Kernel Name"]].replace("b200::", "")}
    for k in KEEP:
        if k in idx:
            e[k] = r[idx[k]]
    out["kernels"].append(e)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(len(out["kernels"]), "kernels ->", sys.argv[2])
