"""CPU-only: how far apart are two valid evaluation orders (fp32 vs the reference's bf16 op order) of the oracle on
the full-width 1-layer checker model of bench.py's parity_full_width leg?  usage: python tools/cpu_fullwidth_inherent.py [scale]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import llava_dpo_oracle as O
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
torch.set_num_threads(os.cpu_count())
cfg = O.OracleConfig(num_layers=1)
p = O.make_params(cfg, seed=0, scale=scale)
batch = O.synthetic_pair_batch(cfg, 1, 48, 64, seed=1234, image_pos=35)
res = {}
with torch.no_grad():
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        t0 = time.time()
        pp = {k: v.to(dt) for k, v in p.items()}
        out = O.policy_logps(pp, cfg, batch["concatenated_input_ids"], batch["concatenated_labels"], batch["images"].to(dt))
        res[name] = (out["logp"].float(), out["per_token_logps"].float(), out["labels"])
        print(name, "logp", out["logp"].float().tolist(), "%.0fs" % (time.time() - t0), flush=True)
mask = res["fp32"][2][:, 1:] != -100
a, b = res["fp32"][1][mask], res["bf16"][1][mask]
print("scale", scale, "summed rel err", float(((res["fp32"][0] - res["bf16"][0]).abs() / res["fp32"][0].abs()).max()),
      "per-token max abs diff", float((a - b).abs().max()), "mean per-token logp", float(a.mean()))
