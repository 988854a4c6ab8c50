"""One DPO step of bench.py's workload between cudaProfilerStart/Stop, for

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py

(the launch list of exactly one timed-region step; shares per kernel via tools/summarize_launches.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from rlaifv_b200.engine import DPOStepEngine
from rlaifv_b200.model import LlavaDims, LlavaDPOPolicy

torch.cuda.set_device(0)
B = bench.PAIRS_PER_GPU
policy = LlavaDPOPolicy(LlavaDims(), torch.device("cuda", 0), seed=0)
policy.stash_act = True            # bench.py's single-GPU policy (SwiGLU product stashed, norms recomputed)
engine = DPOStepEngine(policy, lr=5e-7, weight_decay=0.01, total_steps=2672, micro_pairs=B)
hb = bench.synthetic_batch(0, 0, B)
out = policy.forward_logps(hb["concatenated_input_ids"], hb["concatenated_labels"], hb["images"], keep_stash=False)
hb["ref_win_logp"], hb["ref_rej_logp"] = out["logp"][:B].float().cpu(), out["logp"][B:].float().cpu()
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
for _ in range(2):
    engine.train_step(batch)
engine.opt.wait_all()
torch.cuda.synchronize()
torch.cuda.profiler.start()
engine.train_step(batch)
engine.opt.wait_all()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
