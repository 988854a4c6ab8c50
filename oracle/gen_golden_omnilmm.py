"""ORACLE TOOLING — pins oracle/omnilmm_oracle.py against the UNMODIFIED reference `OmniLMMForCausalLM`
(/root/reference/omnilmm/model/omnilmm.py) driven through the reference's own `forward_DPO` / `dpo_loss`
(muffin/train/trainers.py:66-126) and writes tests/golden/omnilmm/*.npz. Build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_omnilmm.py

`timm` (the EVA-02 tower) and `cv2` are not installed, so they are stubbed at IMPORT time only; the tower is replaced by
a stand-in whose `forward_features` prepends one prefix token to the token sequence it is given. Everything downstream
of the tower — prefix strip, resampler, in-place <im_patch> splice, the HF Mistral decoder (grouped-query attention),
lm_head, get_batch_logps, dpo_loss, and the whole backward — is the reference's own code.
"""
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = "/root/reference"
from oracle import llava_dpo_oracle as O      # noqa: E402
from oracle import omnilmm_oracle as OM       # noqa: E402
from oracle import resampler_oracle as R      # noqa: E402


class StandInTower(nn.Module):
    """What OmniLMMModel touches on the timm model: embed_dim, pos_embed (dtype probe), num_prefix_tokens, blocks,
    attn_pool, forward_features."""

    def __init__(self, embed_dim):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_prefix_tokens = 1
        self.attn_pool = None
        self.pos_embed = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.blocks = nn.ModuleList([nn.Identity(), nn.Identity()])

    def forward_features(self, tokens):
        return torch.cat([torch.zeros_like(tokens[:, :1]), tokens], dim=1)


def import_reference(kv_dim):
    if not os.path.isdir(REF):
        raise SystemExit("reference tree %s not present (golden generation runs in the build container only)" % REF)
    sys.dont_write_bytecode = True
    import transformers                                   # noqa: F401  (its availability probes must run before the stubs)
    from transformers import MistralForCausalLM           # noqa: F401
    timm = types.ModuleType("timm")
    timm.__spec__ = importlib.machinery.ModuleSpec("timm", None)
    timm.models = types.ModuleType("timm.models")
    timm.models.VisionTransformer = type("VisionTransformer", (nn.Module,), {})
    timm.create_model = lambda *a, **k: StandInTower(kv_dim)
    tdata, ttr, tco = (types.ModuleType(n) for n in ("timm.data", "timm.data.transforms", "timm.data.constants"))
    ttr.RandomResizedCropAndInterpolation = object
    tco.IMAGENET_INCEPTION_MEAN = tco.IMAGENET_INCEPTION_STD = (0.5, 0.5, 0.5)
    mods = {"timm": timm, "timm.models": timm.models, "timm.data": tdata, "timm.data.transforms": ttr,
            "timm.data.constants": tco}
    for name in ("cv2", "matplotlib", "matplotlib.pyplot"):
        try:
            __import__(name)
        except Exception:
            mods[name] = types.ModuleType(name)
    sys.modules.update(mods)
    sys.path.insert(0, REF)
    from omnilmm.model.omnilmm import OmniLMMForCausalLM, OmniLMMConfig
    from muffin.train.trainers import forward_DPO, dpo_loss, compute_weighted_logp
    from muffin.eval.muffin_inference_logp import get_batch_logps
    return OmniLMMForCausalLM, OmniLMMConfig, forward_DPO, dpo_loss, get_batch_logps, compute_weighted_logp


def sample(t, n=64):
    f = t.detach().flatten()
    return f[torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()].numpy()


def main():
    dec, res, tok = OM.TINY_OMNI_DEC, OM.TINY_OMNI_RES, OM.TINY_OMNI_TOK
    Model, Config, forward_DPO, dpo_loss, get_batch_logps, compute_weighted_logp = import_reference(res.kv_dim)
    cfg = Config(vocab_size=dec.vocab_size, hidden_size=dec.hidden_size, intermediate_size=dec.intermediate_size,
                 num_hidden_layers=dec.num_layers, num_attention_heads=dec.num_heads,
                 num_key_value_heads=dec.kv_heads, rms_norm_eps=dec.rms_eps, max_position_embeddings=4096,
                 sliding_window=4096, attn_implementation="eager", tie_word_embeddings=False, pad_token_id=0,
                 bos_token_id=1, eos_token_id=2)
    cfg.mm_vision_tower = "stand-in"
    cfg.num_query = res.num_queries
    cfg.image_size = 448
    model = Model(cfg).float()
    vc = model.model.vision_config
    vc.im_patch_token, vc.im_start_token, vc.im_end_token, vc.use_im_start_end = tok.im_patch, tok.im_start, tok.im_end, True
    out_dir = os.path.join(REPO, "tests", "golden", "omnilmm")
    for name, B, seed, ragged, weighted in (("omni_ragged_b2", 2, 41, True, False), ("omni_equal_b1", 1, 42, False, False),
                                            ("omni_weighted_b2", 2, 43, True, True)):
        p = OM.make_omnilmm_params(dec, res, seed=seed)
        missing, unexpected = model.load_state_dict({k: v.clone() for k, v in p.items()}, strict=False)
        missing = [m for m in missing if "vision_tower" not in m and "rotary" not in m]
        assert not missing and not unexpected, (missing, unexpected)
        assert torch.equal(model.model.resampler.pos_embed.data, p["model.resampler.pos_embed"])
        batch = OM.synthetic_omni_batch(dec, res, tok, B, 28, 20, seed=seed + 7, ragged=ragged)
        ids, labels = batch["concatenated_input_ids"], batch["concatenated_labels"]
        vt = batch["vision_tokens"].clone().requires_grad_(True)
        model.train()
        model.zero_grad()
        images = torch.cat([vt, vt], dim=0)                                   # trainers.py:190
        tw = None
        if weighted:
            # --dpo_token_weighted (trainers.py:246-261): per-token log-probs out of forward_DPO, re-weighted by the
            # collator's token weights (1 / mod_token_weight=3 pattern) through the reference's compute_weighted_logp
            g = torch.Generator().manual_seed(seed + 11)
            tw = torch.where(torch.rand(ids.shape[0], ids.shape[1] - 1, generator=g) < 0.3, 3.0, 1.0)
            per_tok = forward_DPO(model, ids, labels.clone(), None, images, token_weighted=True)
            logp = compute_weighted_logp(per_tok, labels, tw, False)
        else:
            logp = forward_DPO(model, ids, labels.clone(), None, images)      # attention_mask=None (trainers.py:199)
            logits = model(input_ids=ids, images=images, attention_mask=None).logits
            per_tok = get_batch_logps(logits, labels.clone(), return_per_token_logp=True)
        ref_w = logp[:B].detach() + 0.3
        ref_r = logp[B:].detach() - 0.2
        losses, cr, rr = dpo_loss(logp[:B], logp[B:], ref_w, ref_r, beta=0.1)
        losses.mean().backward()
        # ---- restatement on the same inputs ----
        po = {k: v.clone().requires_grad_("pos_embed" not in k) for k, v in p.items()}
        vo = batch["vision_tokens"].clone().requires_grad_(True)
        ob = dict(batch, vision_tokens=vo, ref_win_logp=ref_w, ref_rej_logp=ref_r)
        if weighted:
            ob["token_weight"] = tw
        o = OM.omnilmm_dpo_step(po, dec, res, tok, ob)
        o["loss"].backward()
        rel = lambda a, b: float((a.detach() - b.detach()).abs().max() / (b.detach().abs().max() + 1e-30))
        worst = max(rel(o["logp"], logp), rel(o["per_token_logps"], per_tok), rel(o["losses"], losses),
                    rel(vo.grad, vt.grad))
        fx = {"case": name, "B": B, "seed": seed, "ragged": ragged, "logp": logp.detach().numpy(),
              **({"token_weight": tw.numpy()} if weighted else {}),
              "per_token_logps": per_tok.detach().numpy(), "losses": losses.detach().numpy(),
              "chosen_rewards": cr.numpy(), "rejected_rewards": rr.numpy(), "ref_win_logp": ref_w.numpy(),
              "ref_rej_logp": ref_r.numpy(), "loss": float(losses.mean()),
              "dvision_sample": sample(vt.grad, 256), "dvision_norm": float(vt.grad.norm())}
        named = dict(model.named_parameters())
        for k in p:
            if "pos_embed" in k:
                continue
            g_ref, g_o = named[k].grad, po[k].grad
            worst = max(worst, rel(g_o, g_ref))
            fx["gradsample:" + k] = sample(g_ref)
            fx["gradnorm:" + k] = float(g_ref.norm())
        print(f"{name}: restatement vs reference OmniLMMForCausalLM, worst relative error {worst:.2e}")
        assert worst < 5e-5, worst
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **fx)


if __name__ == "__main__":
    main()
