"""ORACLE TOOLING — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

For every case it (1) builds a random-init tiny LlavaLlamaForCausalLM + CLIPVisionTower +
mlp2x_gelu projector exactly as SURVEY.md Appendix A describes, with weights from
oracle.llava_dpo_oracle.make_params (deterministic, so no weight file is committed),
(2) runs the reference's own DataCollatorForDPODataset -> get_beta_and_logps(is_llava15=True) ->
dpo_loss -> loss.backward() on torch-CPU fp32, (3) checks the restated oracle against those
outputs (<= 2e-5 relative; index work bit-exact) and (4) writes inputs + reference outputs as a
small fixture.  The fixture therefore pins BOTH the oracle and (through the oracle-independent
reference outputs it stores) the CUDA path.
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = "/root/reference"


def import_reference():
    """The reference's own modules, imported from /root/reference itself (golden generation never goes through the
    staged copy)."""
    if not os.path.isdir(REF):
        raise SystemExit("reference tree %s not present (golden generation runs in the build container only)" % REF)
    from oracle import stage_ref
    return stage_ref.import_reference(REF)


def build_reference_model(R, cfg, params, load=True):
    from oracle import stage_ref
    return stage_ref.build_reference_model(R, cfg, params, load=load)


def make_instances(batch, ref, B):
    """(rej_dict, win_dict) tuples as DPODataset.__getitem__ yields them
    (muffin/train/train_llava15.py:140-145, muffin/train/train_utils.py:256-262)."""
    ids, labs = batch["concatenated_input_ids"], batch["concatenated_labels"]
    inst = []
    for i in range(B):
        def one(row, kind):
            n = int((ids[row] != 0).sum())     # strip collator padding (pad id 0 never occurs inside)
            d = {"input_ids": ids[row, :n].clone(), "labels": labs[row, :n].clone(), "image": batch["images"][i]}
            d[f"ref_{kind}_logp"] = float(ref[f"ref_{kind}_logp"][i])
            d[f"ref_{kind}_avg_logp"] = float(ref[f"ref_{kind}_avg_logp"][i])
            d[f"ref_{kind}_per_token_logp"] = [0.0] * (n + 600)
            return d
        inst.append((one(B + i, "rej"), one(i, "win")))
    return inst


def reference_bf16_run(R, model, data, Args):
    """The UNMODIFIED reference evaluated the way the shipped recipe runs it: `model.bfloat16()` (DeepSpeed bf16,
    script/zero2.json:2-4), float inputs cast to bf16 (HF Trainer._prepare_inputs under DeepSpeed-bf16,
    HF:trainer.py:2169-2176) and the logits upcast to fp32 before the log-softmax (pinned transformers==4.35.0,
    HF:llama/modeling_llama.py `logits = logits.float()`; 5.5.0 dropped that line, so the upcast is re-applied on
    the forward's output here — the model code itself is untouched).  Returns the log-probs / losses / gradients
    north_star's "1e-3 relative in bf16 vs the reference HF path" refers to."""
    import copy
    mb = copy.deepcopy(model).bfloat16()
    # Environment drift, not reference code: transformers 5.5 keeps RoPE's `inv_freq` as a module buffer, so
    # `.bfloat16()` rounds it and every angle changes; the pinned 4.35.0 builds its cos/sin cache from the fp32
    # inv_freq at construction and only the cached cos/sin are rounded to bf16 (HF:llama/modeling_llama.py
    # LlamaRotaryEmbedding._set_cos_sin_cache).  Restore the fp32 inv_freq to get the pinned version's angles.
    for mod in mb.modules():
        if hasattr(mod, "inv_freq") and hasattr(mod, "original_inv_freq"):
            hd = mb.config.hidden_size // mb.config.num_attention_heads
            inv = 1.0 / (mb.config.rope_theta if hasattr(mb.config, "rope_theta") and mb.config.rope_theta else 10000.0) ** (
                torch.arange(0, hd, 2, dtype=torch.int64).float() / hd)
            mod.inv_freq = inv
            mod.original_inv_freq = inv.clone()
    mb.train()
    mb.zero_grad(set_to_none=True)
    inner_forward = mb.forward

    def forward_upcast(*a, **kw):
        out = inner_forward(*a, **kw)
        out.logits = out.logits.float()
        return out

    mb.forward = forward_upcast
    d = dict(data)
    d["images"] = data["images"].to(torch.bfloat16)
    with torch.no_grad():
        imgs2 = torch.cat([d["images"], d["images"]], 0)
        _, _, _, _, emb, new_labels = mb.prepare_inputs_labels_for_multimodal(
            input_ids=data["concatenated_input_ids"].clone(), position_ids=None, attention_mask=None,
            past_key_values=None, labels=data["concatenated_labels"].clone(), images=imgs2)
        logits = mb.forward(inputs_embeds=emb, labels=None).logits
        per_tok, _, _ = R["get_batch_logps"](logits, new_labels, return_all=True)
        feats = mb.get_model().get_vision_tower()(d["images"])
        proj = mb.get_model().mm_projector(feats)
    pw, pr, rw, rr, beta = R["get_beta_and_logps"](d, mb, Args(), is_llava15=True)
    losses, cr, rj = R["dpo_loss"](pw, pr, rw, rr, beta=beta)
    loss = losses.mean()
    loss.backward()
    grads = {k: v.grad.detach().float().clone() for k, v in mb.named_parameters() if v.grad is not None}
    return dict(per_token_logps=per_tok.float(), policy_win_logp=pw.detach().float(), policy_rej_logp=pr.detach().float(),
                losses=losses.detach().float(), chosen_rewards=cr.float(), rejected_rewards=rj.float(),
                loss=float(loss), grads=grads, clip_features=feats.float(), projected_rows=proj.float())


CASES = {
    # name: (B, prompt_len, resp_len, seed, image_pos, ragged, param_scale)
    "tiny_equal": (1, 20, 16, 11, None, False, 1.0),
    "tiny_ragged_b2": (2, 24, 20, 12, 5, True, 1.0),
    "tiny_image_first": (2, 16, 12, 13, 1, True, 1.0),
    # cooler weights: logits of the magnitude a real checkpoint produces, so the strict 1e-3 bf16 gate applies
    "cool_ragged_b2": (2, 24, 40, 14, 9, True, 0.4),
    "cool_long_b1": (1, 30, 150, 15, 20, False, 0.4),
    # grouped-query decoder (4 query / 2 kv heads): reference = the same LlavaLlamaForCausalLM with
    # num_key_value_heads=2 (HF Llama's repeat_kv path, the arithmetic of the Mistral LLM in OmniLMM-12B)
    "gqa_ragged_b2": (2, 24, 40, 16, 9, True, 0.4, "TINY_GQA"),
}


def main():
    from oracle import llava_dpo_oracle as O
    R = import_reference()
    cfg = O.TINY
    models = {}
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    class Tok:
        pad_token_id = 0

    class Args:
        dpo_use_average = False
        dpo_token_weighted = False
        task = "DPO"

    for name, case in CASES.items():
        B, P, Rl, seed, ipos, ragged, pscale = case[:7]
        cfg_name = case[7] if len(case) > 7 else "TINY"
        cfg = O.CONFIGS[cfg_name]
        if (pscale, cfg_name) not in models:
            prm = O.make_params(cfg, seed=0, scale=pscale)
            mdl = build_reference_model(R, cfg, prm)
            mdl.train()
            models[(pscale, cfg_name)] = (prm, mdl)
        params, model = models[(pscale, cfg_name)]
        batch = O.synthetic_pair_batch(cfg, B, P, Rl, seed, image_pos=ipos, ragged=ragged)
        g = torch.Generator().manual_seed(seed + 100)
        ref = {k: (-40.0 + 3.0 * torch.randn(B, generator=g)) for k in ("ref_win_logp", "ref_rej_logp")}
        ref["ref_win_avg_logp"] = ref["ref_win_logp"] / Rl
        ref["ref_rej_avg_logp"] = ref["ref_rej_logp"] / Rl
        collator = R["DataCollatorForDPODataset"](tokenizer=Tok(), beta=0.1, mod_token_weight=1.0)
        data = collator(make_instances(batch, ref, B))
        assert torch.equal(data["concatenated_input_ids"], batch["concatenated_input_ids"])
        assert torch.equal(data["concatenated_labels"], batch["concatenated_labels"])
        # --- reference forward/backward (unmodified functions) ---
        model.zero_grad(set_to_none=True)
        keep_ids = data["concatenated_input_ids"].clone()
        keep_labels = data["concatenated_labels"].clone()
        # also capture the spliced embeds/labels the reference builds
        with torch.no_grad():
            _, _, _, _, ref_embeds, ref_new_labels = model.prepare_inputs_labels_for_multimodal(
                input_ids=keep_ids, position_ids=None, attention_mask=None, past_key_values=None,
                labels=keep_labels, images=torch.cat([data["images"], data["images"]], 0))
            ref_logits = model.forward(inputs_embeds=ref_embeds, labels=None).logits.float()
            ref_per_tok, _, _ = R["get_batch_logps"](ref_logits, ref_new_labels, return_all=True)
            ref_feats = model.get_model().get_vision_tower()(data["images"])        # clip_encoder.py:46-58
            ref_proj = model.get_model().mm_projector(ref_feats)                    # llava_arch.py:147
        pw, pr, rw, rr, beta = R["get_beta_and_logps"](dict(data), model, Args(), is_llava15=True)
        losses, cr, rj = R["dpo_loss"](pw, pr, rw, rr, beta=beta)
        loss = losses.mean()
        loss.backward()
        ref_grads = {k: v.grad.detach().clone() for k, v in model.named_parameters() if v.grad is not None}
        assert all("vision_tower" not in k for k in ref_grads), "CLIP must stay frozen"
        model.zero_grad(set_to_none=True)
        rb = reference_bf16_run(R, model, data, Args)       # the same unmodified reference, bf16 as shipped
        # --- oracle ---
        op = {k: v.clone().requires_grad_(k.startswith(O.TRAINABLE_PREFIXES)) for k, v in params.items()}
        ob = dict(batch, ref_win_logp=ref["ref_win_logp"], ref_rej_logp=ref["ref_rej_logp"])
        oo = O.dpo_step(op, cfg, ob, beta=0.1)
        oo["loss"].backward()

        def rel(a, b):
            return float((a - b).abs().max() / (b.abs().max() + 1e-12))

        assert torch.equal(oo["labels"], ref_new_labels), "splice labels not bit-exact"
        assert torch.equal(oo["inputs_embeds"].detach(), ref_embeds), "splice rows not bit-exact"
        checks = {
            "per_token_logps": rel(oo["per_token_logps"].detach(), ref_per_tok),
            "policy_win_logp": rel(oo["policy_win_logp"].detach(), pw.detach()),
            "policy_rej_logp": rel(oo["policy_rej_logp"].detach(), pr.detach()),
            "losses": rel(oo["losses"].detach(), losses.detach()),
            "chosen_rewards": rel(oo["chosen_rewards"], cr),
        }
        gsel = ["model.embed_tokens.weight", "lm_head.weight", "model.mm_projector.0.weight",
                "model.mm_projector.2.bias", "model.layers.0.self_attn.q_proj.weight",
                "model.layers.1.mlp.down_proj.weight", "model.layers.0.input_layernorm.weight",
                "model.norm.weight"]
        for k in ref_grads:
            checks["grad:" + k] = rel(op[k].grad, ref_grads[k])
        worst = max(checks.values())
        print(f"[{name}] oracle vs reference: worst rel err {worst:.3e} over {len(checks)} quantities; "
              f"loss={float(loss):.6f} T={ref_new_labels.shape[1]}")
        assert worst < 2e-5, {k: v for k, v in checks.items() if v >= 2e-5}
        fx = dict(
            B=np.int64(B), prompt_len=np.int64(P), resp_len=np.int64(Rl), seed=np.int64(seed),
            image_pos=np.int64(-1 if ipos is None else ipos), ragged=np.int64(int(ragged)),
            params_checksum=np.float64(O.params_checksum(params)), param_scale=np.float64(pscale),
            cfg_name=np.array(cfg_name),
            concatenated_input_ids=keep_ids.numpy(), concatenated_labels=keep_labels.numpy(),
            images=data["images"].numpy().astype(np.float32),
            ref_win_logp=ref["ref_win_logp"].numpy(), ref_rej_logp=ref["ref_rej_logp"].numpy(),
            beta=np.float64(0.1),
            spliced_labels=ref_new_labels.numpy(),
            spliced_embeds_rowsum=ref_embeds.double().sum(-1).numpy(),
            per_token_logps=ref_per_tok.numpy(), policy_win_logp=pw.detach().numpy(),
            policy_rej_logp=pr.detach().numpy(), losses=losses.detach().numpy(),
            chosen_rewards=cr.numpy(), rejected_rewards=rj.numpy(), loss=np.float64(float(loss)),
            clip_features=ref_feats.numpy(), projected_rows=ref_proj.detach().numpy(),
            # the reference run in bf16 (model.bfloat16(), bf16 images, fp32 logits) — reference_bf16_run
            bf16_per_token_logps=rb["per_token_logps"].numpy(), bf16_policy_win_logp=rb["policy_win_logp"].numpy(),
            bf16_policy_rej_logp=rb["policy_rej_logp"].numpy(), bf16_losses=rb["losses"].numpy(),
            bf16_chosen_rewards=rb["chosen_rewards"].numpy(), bf16_rejected_rewards=rb["rejected_rewards"].numpy(),
            bf16_loss=np.float64(rb["loss"]), bf16_clip_features=rb["clip_features"].numpy(),
            bf16_projected_rows=rb["projected_rows"].numpy(),
        )
        ref32 = torch.cat([pw.detach(), pr.detach()])
        refbf = torch.cat([rb["policy_win_logp"], rb["policy_rej_logp"]])
        print(f"    reference bf16 vs fp32: summed logp {rel(refbf, ref32):.2e}, loss {abs(rb['loss'] - float(loss)) / abs(float(loss)):.2e}")
        for k in gsel:
            gk = ref_grads[k]
            fx["gradnorm:" + k] = np.float64(float(gk.double().norm()))
            flat = gk.flatten()
            idx = torch.linspace(0, flat.numel() - 1, 64).long()
            fx["gradsample:" + k] = flat[idx].numpy()
            gb = rb["grads"][k]
            fx["bf16_gradnorm:" + k] = np.float64(float(gb.double().norm()))
            fx["bf16_gradsample:" + k] = gb.flatten()[idx].numpy()
            fx["bf16_gradrelerr:" + k] = np.float64(float((gb.double() - gk.double()).norm() / gk.double().norm()))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **fx)
    # ---- collator fixture: the reference's DataCollatorForDPODataset on pairs that share edits ----
    g = torch.Generator().manual_seed(77)
    inst, flat = [], {}
    for i in range(3):
        base = torch.randint(3, 500, (40 + 7 * i,), generator=g)
        win = base.clone()
        rej = base.clone()
        rej[10:14] = torch.randint(3, 500, (4,), generator=g)              # substitution
        rej = torch.cat([rej[:25], torch.randint(3, 500, (3 + i,), generator=g), rej[25:]])   # insertion
        win = torch.cat([win[:33], win[36:]])                              # deletion on the win side
        prompt = torch.randint(3, 500, (12,), generator=g)
        prompt[0] = 1
        prompt[5] = -200
        def mk(resp, kind):
            ids = torch.cat([prompt, resp])
            labs = torch.cat([torch.full((12,), -100), resp])
            d = {"input_ids": ids, "labels": labs, "image": torch.randn(3, 8, 8, generator=g),
                 f"ref_{kind}_logp": float(-10.0 - i), f"ref_{kind}_avg_logp": float(-0.5 - 0.1 * i),
                 f"ref_{kind}_per_token_logp": torch.randn(len(ids) + 5, generator=g).tolist()}
            return d
        r, w = mk(rej, "rej"), mk(win, "win")
        inst.append((r, w))
        for kind, d in (("rej", r), ("win", w)):
            flat[f"in{i}_{kind}_input_ids"] = d["input_ids"].numpy()
            flat[f"in{i}_{kind}_labels"] = d["labels"].numpy()
            flat[f"in{i}_{kind}_image"] = d["image"].numpy()
            flat[f"in{i}_{kind}_logp"] = np.float64(d[f"ref_{kind}_logp"])
            flat[f"in{i}_{kind}_avg_logp"] = np.float64(d[f"ref_{kind}_avg_logp"])
            flat[f"in{i}_{kind}_per_token"] = np.asarray(d[f"ref_{kind}_per_token_logp"], dtype=np.float64)
    coll = R["DataCollatorForDPODataset"](tokenizer=Tok(), beta=0.1, mod_token_weight=3.0)
    outb = coll(inst)
    assert len(outb) == 20
    for k, v in outb.items():
        flat["out_" + k] = v.numpy() if torch.is_tensor(v) else np.float64(v)
    os.makedirs(os.path.join(REPO, "tests", "golden_host"), exist_ok=True)
    np.savez_compressed(os.path.join(REPO, "tests", "golden_host", "collator_case.npz"), n=np.int64(3), **flat)
    # ---- sample-encoding fixture: reference preprocess_v1 / encode_multimodal_preference_sample ----
    from functools import partial
    from oracle.toy_tokenizer import ToyTokenizer
    from muffin.train.train_utils import encode_multimodal_preference_sample as ref_encode, preprocess_v1 as ref_pre
    tok = ToyTokenizer()
    enc = {}
    samples = [("<image>\nwhat is shown here ?", "a red bus on a street", "a blue car"),
               ("describe the <image> scene in detail please", "two dogs play . they run", "one cat sleeps"),
               ("<image>\nhow many ?", "three", "four apples on the table near a window")]
    for i, (q, c, r) in enumerate(samples):
        src = {"question": {"from": "human", "value": q}, "chosen": {"from": "gpt", "value": c},
               "rejected": {"from": "gpt", "value": r}, "image": "IMG"}
        cfgm = {"image_processor": lambda im: torch.zeros(3, 4, 4), "is_multimodal": True, "image_token_len": 576,
                "use_im_start_end": False, "keep_image_tag": True}
        rej_d, win_d = ref_encode(src, tok, cfgm, preprocess_func=partial(ref_pre, has_image=True))
        enc[f"s{i}_q"], enc[f"s{i}_c"], enc[f"s{i}_r"] = np.array(q), np.array(c), np.array(r)
        enc[f"s{i}_win_ids"], enc[f"s{i}_win_labels"] = win_d["input_ids"].numpy(), win_d["labels"].numpy()
        enc[f"s{i}_rej_ids"], enc[f"s{i}_rej_labels"] = rej_d["input_ids"].numpy(), rej_d["labels"].numpy()
    np.savez_compressed(os.path.join(REPO, "tests", "golden_host", "encode_case.npz"), n=np.int64(len(samples)), **enc)
    print("golden fixtures written to", out_dir)


if __name__ == "__main__":
    main()
