"""Host-side cost per C-ABI launch: tiny model (GPU work negligible), many steps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import llava_dpo_oracle as O
from rlaifv_b200 import lib
from rlaifv_b200.engine import DPOStepEngine
from rlaifv_b200.model import LlavaDims, LlavaDPOPolicy
c = O.TINY
dims = LlavaDims(vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                 num_layers=8, num_heads=c.num_heads, clip_hidden=c.clip_hidden,
                 clip_intermediate=c.clip_intermediate, clip_layers=c.clip_layers, clip_heads=c.clip_heads,
                 image_size=c.image_size, patch_size=c.patch_size)
pol = LlavaDPOPolicy(dims, "cuda", seed=0)
eng = DPOStepEngine(pol, lr=1e-4, total_steps=100, constant_lr=True)
batch = O.synthetic_pair_batch(c, 2, 24, 20, seed=5, image_pos=7)
batch["ref_win_logp"] = torch.tensor([-60.0, -70.0]); batch["ref_rej_logp"] = torch.tensor([-61.0, -69.0]); batch["beta"] = 0.1
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
for _ in range(3): eng.train_step(batch)
torch.cuda.synchronize()
n0 = lib.launch_count(); t0 = time.perf_counter()
for _ in range(20): eng.train_step(batch)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
n = lib.launch_count() - n0
print(f"launches/step {n/20:.0f}; host enqueue {1e6*(t1-t0)/n:.1f} us per launch; incl. drain {1e6*(t2-t0)/n:.1f} us")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): eng.train_step(batch)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
