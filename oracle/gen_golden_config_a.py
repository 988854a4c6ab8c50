"""ORACLE TOOLING — BASELINE.json configs[0] at FULL WIDTH and FULL DEPTH, run through the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_config_a.py        (build container only; ~10 min, ~45 GB RAM)

LLaVA-1.5-7B dimensions (h=4096, ffn=11008, 32 decoder layers, vocab 32000) + CLIP-ViT-L/14-336 (24 layers, layer -2
selected), 1 synthetic pair, 336 px image, 48-token prompt with the image slot at 35, 64-token chosen / rejected
responses (T = 687) — SURVEY.md §8c "memory note": forward-only fits the build container's 62 GB.

Weights: oracle.llava_dpo_oracle.make_params(OracleConfig(), seed=0, scale=CHECKER_PARAM_SCALE) — deterministic, so
only the reference's OUTPUTS are committed (tests/golden_full/config_a.npz, a few KB): per-token / summed log-probs
and DPO losses of
  (1) the reference in fp32 (get_beta_and_logps under no_grad -> dpo_loss), and
  (2) the reference as shipped: model.bfloat16(), bf16 images, fp32 logits (gen_golden.reference_bf16_run's recipe).
tests/test_gpu_config_a.py regenerates the same weights on the GPU box and holds the CUDA path to (2) at 1e-3
(north_star) and reports (1) beside it.
"""
import gc
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CHECKER_PARAM_SCALE = 0.3   # same conditioning argument as bench.CHECKER_PARAM_SCALE: logits of a real checkpoint's
                            # magnitude (mean per-token log-prob ~ -ln 32000) instead of an ill-conditioned random net


def main():
    from oracle import gen_golden as G
    from oracle import llava_dpo_oracle as O
    R = G.import_reference()
    cfg = O.OracleConfig()
    t0 = time.time()
    params = O.make_params(cfg, seed=0, scale=CHECKER_PARAM_SCALE)
    checksum = O.params_checksum(params)
    print("params generated in %.0fs, checksum %.6e" % (time.time() - t0, checksum), flush=True)
    batch = O.synthetic_pair_batch(cfg, 1, 48, 64, seed=1234, image_pos=35)
    g = torch.Generator().manual_seed(99)
    ref = {"ref_win_logp": torch.tensor([-700.0]), "ref_rej_logp": torch.tensor([-690.5])}
    ref["ref_win_avg_logp"] = ref["ref_win_logp"] / 64
    ref["ref_rej_avg_logp"] = ref["ref_rej_logp"] / 64

    class Tok:
        pad_token_id = 0

    class Args:
        dpo_use_average = False
        dpo_token_weighted = False
        task = "DPO"

    collator = R["DataCollatorForDPODataset"](tokenizer=Tok(), beta=0.1, mod_token_weight=1.0)
    data = collator(G.make_instances(batch, ref, 1))
    assert torch.equal(data["concatenated_input_ids"], batch["concatenated_input_ids"])

    # model on the meta device, parameters assigned (no second 27 GB copy)
    with torch.device("meta"):
        model = G.build_reference_model(R, cfg, None, load=False)
    missing, unexpected = model.load_state_dict(params, strict=False, assign=True)
    missing = [m for m in missing if "rotary" not in m and "position_ids" not in m]
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    for mod in model.modules():      # non-persistent buffers (RoPE inv_freq, CLIP position_ids) are still meta: rebuild
        if hasattr(mod, "inv_freq") and mod.inv_freq.is_meta:
            hd = cfg.head_dim
            inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
            mod.inv_freq = inv
            if hasattr(mod, "original_inv_freq"):
                mod.original_inv_freq = inv.clone()
        if hasattr(mod, "position_ids") and torch.is_tensor(mod.position_ids) and mod.position_ids.is_meta:
            mod.position_ids = torch.arange(mod.position_ids.shape[-1]).expand(1, -1)
    model.eval()
    del params
    gc.collect()

    def run(m, d, upcast):
        inner = m.forward

        def fwd(*a, **kw):
            out = inner(*a, **kw)
            out.logits = out.logits.float()
            return out
        if upcast:
            m.forward = fwd
        with torch.no_grad():
            imgs2 = torch.cat([d["images"], d["images"]], 0)
            _, _, _, _, emb, new_labels = m.prepare_inputs_labels_for_multimodal(
                input_ids=d["concatenated_input_ids"].clone(), position_ids=None, attention_mask=None,
                past_key_values=None, labels=d["concatenated_labels"].clone(), images=imgs2)
            logits = m.forward(inputs_embeds=emb, labels=None).logits.float()
            per_tok, _, _ = R["get_batch_logps"](logits, new_labels, return_all=True)
            del logits
            pw, pr, rw, rr, beta = R["get_beta_and_logps"](dict(d), m, Args(), is_llava15=True)
            losses, cr, rj = R["dpo_loss"](pw, pr, rw, rr, beta=beta)
        if upcast:
            m.forward = inner
        return dict(per_tok=per_tok.float(), pw=pw.float(), pr=pr.float(), losses=losses.float(), cr=cr.float(),
                    rj=rj.float(), labels=new_labels)

    t0 = time.time()
    r32 = run(model, data, False)
    print("fp32 reference forward x2: %.0fs; logp %s %s loss %.6f" % (time.time() - t0, r32["pw"].tolist(),
                                                                      r32["pr"].tolist(), float(r32["losses"].mean())),
          flush=True)
    model.bfloat16()
    for mod in model.modules():      # pinned 4.35.0 angles: fp32 inv_freq, bf16 cos/sin (see gen_golden.reference_bf16_run)
        if hasattr(mod, "inv_freq"):
            hd = cfg.head_dim
            inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
            mod.inv_freq = inv
            if hasattr(mod, "original_inv_freq"):
                mod.original_inv_freq = inv.clone()
    gc.collect()
    db = dict(data)
    db["images"] = data["images"].to(torch.bfloat16)
    t0 = time.time()
    rb = run(model, db, True)
    print("bf16 reference forward x2: %.0fs; logp %s %s loss %.6f" % (time.time() - t0, rb["pw"].tolist(),
                                                                      rb["pr"].tolist(), float(rb["losses"].mean())),
          flush=True)
    out_dir = os.path.join(REPO, "tests", "golden_full")
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(
        os.path.join(out_dir, "config_a.npz"),
        param_scale=np.float64(CHECKER_PARAM_SCALE), params_checksum=np.float64(checksum),
        prompt_len=np.int64(48), resp_len=np.int64(64), seed=np.int64(1234), image_pos=np.int64(35),
        concatenated_input_ids=data["concatenated_input_ids"].numpy(),
        concatenated_labels=data["concatenated_labels"].numpy(),
        images_checksum=np.float64(float(data["images"].double().abs().sum())),
        ref_win_logp=ref["ref_win_logp"].numpy(), ref_rej_logp=ref["ref_rej_logp"].numpy(), beta=np.float64(0.1),
        spliced_labels=r32["labels"].numpy(),
        per_token_logps=r32["per_tok"].numpy(), policy_win_logp=r32["pw"].numpy(), policy_rej_logp=r32["pr"].numpy(),
        losses=r32["losses"].numpy(), chosen_rewards=r32["cr"].numpy(), rejected_rewards=r32["rj"].numpy(),
        bf16_per_token_logps=rb["per_tok"].numpy(), bf16_policy_win_logp=rb["pw"].numpy(),
        bf16_policy_rej_logp=rb["pr"].numpy(), bf16_losses=rb["losses"].numpy(),
        bf16_chosen_rewards=rb["cr"].numpy(), bf16_rejected_rewards=rb["rj"].numpy())
    print("written", os.path.join(out_dir, "config_a.npz"))


if __name__ == "__main__":
    main()
