"""Tiny full DPO step (full FT and LoRA) for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import llava_dpo_oracle as O
from rlaifv_b200.engine import DPOStepEngine
from rlaifv_b200.model import LlavaDims, LlavaDPOPolicy
c = O.TINY
dims = LlavaDims(vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                 num_layers=c.num_layers, num_heads=c.num_heads, clip_hidden=c.clip_hidden,
                 clip_intermediate=c.clip_intermediate, clip_layers=c.clip_layers, clip_heads=c.clip_heads,
                 image_size=c.image_size, patch_size=c.patch_size)
for lora in (False, True):
    pol = LlavaDPOPolicy(dims, "cuda", hf_state=O.make_params(c, seed=0, scale=0.4))
    if lora:
        pol.enable_lora(r=8, alpha=2.0, init_b_zero=False)
    eng = DPOStepEngine(pol, lr=1e-4, total_steps=10, constant_lr=True, micro_pairs=1 if lora else None)
    batch = O.synthetic_pair_batch(c, 2, 24, 150, seed=5, image_pos=7, ragged=True)   # T > 128: several q/kv tiles
    batch["ref_win_logp"] = torch.tensor([-600.0, -700.0]); batch["ref_rej_logp"] = torch.tensor([-610.0, -690.0]); batch["beta"] = 0.1
    m = eng.train_step(batch)
    torch.cuda.synchronize()
    print("lora" if lora else "full", "loss", float(m[0]))
print("SANITIZE_STEP_DONE")
