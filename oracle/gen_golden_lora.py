"""ORACLE TOOLING — pins the LoRA-DPO branch (BASELINE config e) against the UNMODIFIED reference model.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_lora.py            (build container only)

peft (the reference's adapter library, pyproject.toml:20, muffin/train/train_llava15_lora.py:304-318) is not
installed here, but the reference MODEL is, and LoRA at dropout 0 is an exact reparametrisation of it:

    y = W x + s B (A x) = (W + s B A) x = W' x,            s = lora_alpha / r

so running the unmodified `LlavaLlamaForCausalLM` with the merged weights W' through the reference's own
`get_beta_and_logps -> dpo_loss -> backward` gives the adapter run's log-probs and losses exactly, and its parameter
gradient dL/dW' gives the adapter gradients by the chain rule of the same identity:

    dL/dA = s B^T (dL/dW')          dL/dB = s (dL/dW') A^T

The projector (trainable in the LoRA recipe, llava/model/llava_arch.py:88-91) is read off directly.  This script
checks oracle.llava_dpo_oracle's LoRA branch (lora_linear: the unmerged two-GEMM form peft executes) against those
reference results in fp32 and writes tests/golden/lora/*.npz; tests/test_gpu_lora.py holds the CUDA path to them.
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

R, ALPHA = 8, 2.0            # scaling 0.25, the shipped r=64 / alpha=16 ratio (train_llava15_lora.py:113-114)
CASES = {
    # name: (B, prompt_len, resp_len, seed, image_pos, ragged, param_scale, lora_seed, b_std)
    "lora_ragged_b2": (2, 24, 30, 31, 7, True, 0.4, 3, 0.02),
    "lora_equal_b1": (1, 20, 40, 32, 5, False, 0.4, 4, 0.05),
}


def merged_params(O, cfg, params, lora, scaling):
    out = dict(params)
    for i in range(cfg.num_layers):
        for t in O.LORA_TARGETS:
            n = f"model.layers.{i}.{t}"
            out[n + ".weight"] = params[n + ".weight"] + scaling * (lora[n + ".lora_B.weight"] @ lora[n + ".lora_A.weight"])
    return out


def main():
    from oracle import gen_golden as G
    from oracle import llava_dpo_oracle as O
    Rf = G.import_reference()
    cfg = O.TINY
    scaling = ALPHA / R
    out_dir = os.path.join(REPO, "tests", "golden", "lora")
    os.makedirs(out_dir, exist_ok=True)

    class Tok:
        pad_token_id = 0

    class Args:
        dpo_use_average = False
        dpo_token_weighted = False
        task = "DPO"

    def rel(a, b):
        return float((a - b).abs().max() / (b.abs().max() + 1e-12))

    for name, (B, P, Rl, seed, ipos, ragged, pscale, lseed, bstd) in CASES.items():
        params = O.make_params(cfg, seed=0, scale=pscale)
        lora = O.make_lora_params(cfg, r=R, seed=lseed, b_std=bstd)
        model = G.build_reference_model(Rf, cfg, merged_params(O, cfg, params, lora, scaling))
        model.train()
        batch = O.synthetic_pair_batch(cfg, B, P, Rl, seed, image_pos=ipos, ragged=ragged)
        g = torch.Generator().manual_seed(seed + 100)
        ref = {k: (-150.0 + 3.0 * torch.randn(B, generator=g)) for k in ("ref_win_logp", "ref_rej_logp")}
        ref["ref_win_avg_logp"] = ref["ref_win_logp"] / Rl
        ref["ref_rej_avg_logp"] = ref["ref_rej_logp"] / Rl
        data = Rf["DataCollatorForDPODataset"](tokenizer=Tok(), beta=0.1, mod_token_weight=1.0)(
            G.make_instances(batch, ref, B))
        assert torch.equal(data["concatenated_input_ids"], batch["concatenated_input_ids"])
        with torch.no_grad():
            _, _, _, _, emb, new_labels = model.prepare_inputs_labels_for_multimodal(
                input_ids=data["concatenated_input_ids"].clone(), position_ids=None, attention_mask=None,
                past_key_values=None, labels=data["concatenated_labels"].clone(),
                images=torch.cat([data["images"], data["images"]], 0))
            ref_per_tok, _, _ = Rf["get_batch_logps"](model.forward(inputs_embeds=emb, labels=None).logits.float(),
                                                     new_labels, return_all=True)
        model.zero_grad(set_to_none=True)
        pw, pr, rw, rr, beta = Rf["get_beta_and_logps"](dict(data), model, Args(), is_llava15=True)
        losses, cr, rj = Rf["dpo_loss"](pw, pr, rw, rr, beta=beta)
        losses.mean().backward()
        gW = {k: v.grad.detach() for k, v in model.named_parameters() if v.grad is not None}
        ref_grads = {}
        for i in range(cfg.num_layers):
            for t in O.LORA_TARGETS:
                n = f"model.layers.{i}.{t}"
                A, Bm, dW = lora[n + ".lora_A.weight"], lora[n + ".lora_B.weight"], gW[n + ".weight"]
                ref_grads[n + ".lora_A.weight"] = scaling * (Bm.t() @ dW)
                ref_grads[n + ".lora_B.weight"] = scaling * (dW @ A.t())
        for k in ("model.mm_projector.0.weight", "model.mm_projector.0.bias", "model.mm_projector.2.weight",
                  "model.mm_projector.2.bias"):
            ref_grads[k] = gW[k]
        # --- the oracle's unmerged LoRA branch against the reference ---
        op = {k: v.clone().requires_grad_(k.startswith("model.mm_projector.")) for k, v in params.items()}
        op.update({k: v.clone().requires_grad_(True) for k, v in lora.items()})
        oo = O.dpo_step(op, cfg, dict(batch, ref_win_logp=ref["ref_win_logp"], ref_rej_logp=ref["ref_rej_logp"]),
                        beta=0.1)
        oo["loss"].backward()
        assert torch.equal(oo["labels"], new_labels)
        checks = {"per_token_logps": rel(oo["per_token_logps"].detach(), ref_per_tok),
                  "policy_win_logp": rel(oo["policy_win_logp"].detach(), pw.detach()),
                  "policy_rej_logp": rel(oo["policy_rej_logp"].detach(), pr.detach()),
                  "losses": rel(oo["losses"].detach(), losses.detach())}
        for k, v in ref_grads.items():
            checks["grad:" + k] = rel(op[k].grad, v)
        worst = max(checks.values())
        print(f"[{name}] oracle LoRA branch vs merged-weight reference: worst rel err {worst:.3e} over {len(checks)} "
              f"quantities; loss={float(losses.mean()):.6f}")
        assert worst < 5e-5, {k: v for k, v in checks.items() if v >= 5e-5}
        # the adapters must matter (otherwise the fixture pins nothing)
        base = Rf and O.policy_logps(params, cfg, batch["concatenated_input_ids"], batch["concatenated_labels"],
                                     batch["images"])
        assert rel(base["logp"].detach(), torch.cat([pw, pr]).detach()) > 1e-3
        fx = dict(B=np.int64(B), prompt_len=np.int64(P), resp_len=np.int64(Rl), seed=np.int64(seed),
                  image_pos=np.int64(ipos), ragged=np.int64(int(ragged)), param_scale=np.float64(pscale),
                  lora_seed=np.int64(lseed), lora_b_std=np.float64(bstd), r=np.int64(R), alpha=np.float64(ALPHA),
                  params_checksum=np.float64(O.params_checksum(params)), lora_checksum=np.float64(O.params_checksum(lora)),
                  concatenated_input_ids=data["concatenated_input_ids"].numpy(),
                  concatenated_labels=data["concatenated_labels"].numpy(), images=data["images"].numpy(),
                  ref_win_logp=ref["ref_win_logp"].numpy(), ref_rej_logp=ref["ref_rej_logp"].numpy(), beta=np.float64(0.1),
                  spliced_labels=new_labels.numpy(), per_token_logps=ref_per_tok.numpy(),
                  policy_win_logp=pw.detach().numpy(), policy_rej_logp=pr.detach().numpy(),
                  losses=losses.detach().numpy(), chosen_rewards=cr.numpy(), rejected_rewards=rj.numpy(),
                  loss=np.float64(float(losses.mean())))
        for k, v in ref_grads.items():
            fx["gradnorm:" + k] = np.float64(float(v.double().norm()))
            flat = v.flatten()
            idx = torch.linspace(0, flat.numel() - 1, 64).long()
            fx["gradsample:" + k] = flat[idx].numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **fx)
    print("LoRA fixtures written to", out_dir)


if __name__ == "__main__":
    main()
