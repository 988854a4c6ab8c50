"""TEST INFRASTRUCTURE ONLY — CPU restatement of the OmniLMM-12B DPO policy path DOWNSTREAM of the vision tower
(BASELINE config d). Only tests/ may import this.

Follows, from the point where the (timm EVA-02) tower has produced its token sequence:
  OmniLMMModel.get_vision_embedding   /root/reference/omnilmm/model/omnilmm.py:107-120   prefix tokens dropped,
                                                                                         resampler applied
  OmniLMMModel.forward                omnilmm/model/omnilmm.py:183-265   embed_tokens, then the num_query rows after
                                                                         each <im_start> are REPLACED by the image's
                                                                         resampled features (sequence length unchanged);
                                                                         a sequence without <im_patch> consumes no image
  MistralModel / lm_head              omnilmm/model/omnilmm.py:259-265, 318-319 (HF MistralForCausalLM: Llama block
                                      arithmetic with grouped-query attention; the 4096-token sliding window is inert
                                      at the <= 2048-token sequences of this path)
  forward_DPO + get_batch_logps       muffin/train/trainers.py:66-88, muffin/eval/muffin_inference_logp.py:82-115
  dpo_loss                            muffin/train/trainers.py:91-126
The vision tower itself is NOT restated: timm (==0.9.10) is absent and the reference only names the architecture
(`timm.create_model('eva02_enormous_patch14_clip_224.laion2b_plus')`, omnilmm.py:31-36), so there is nothing to
pin it against; the boundary of this oracle (and of the CUDA module) is the tower's output tokens.

PINNED: oracle/gen_golden_omnilmm.py runs the UNMODIFIED reference `OmniLMMForCausalLM` (with a stand-in tower that
returns the given tokens — `timm` is stubbed at import time only) through the reference's own `forward_DPO` and
`dpo_loss`, and writes tests/golden/omnilmm/*.npz.
"""
from dataclasses import dataclass

import torch

from . import llava_dpo_oracle as O
from . import resampler_oracle as R


@dataclass(frozen=True)
class OmniTokens:
    im_patch: int
    im_start: int
    im_end: int


# tiny Mistral-style decoder: 4 query heads / 2 kv heads of width 128 (the CUDA attention kernels' head_dim),
# 16 learned queries over a 12x12 vision-token grid of width 192
TINY_OMNI_DEC = O.OracleConfig(vocab_size=512, hidden_size=512, intermediate_size=768, num_layers=2, num_heads=4,
                               num_kv_heads=2, rms_eps=1e-5)
TINY_OMNI_RES = R.ResamplerConfig(grid_size=4, embed_dim=512, num_heads=4, kv_dim=192, kv_tokens=144)
TINY_OMNI_TOK = OmniTokens(im_patch=500, im_start=501, im_end=502)


def make_omnilmm_params(dec_cfg, res_cfg, seed=0, scale=0.4):
    """HF-named state: Mistral decoder (`model.*`, `lm_head.weight`) + `model.resampler.*`."""
    full = O.make_params(dec_cfg, seed=seed, scale=scale)
    p = {k: v for k, v in full.items() if "vision_tower" not in k and "mm_projector" not in k}
    for k, v in R.make_resampler_params(res_cfg, seed=seed + 1).items():
        p["model.resampler." + k] = v
    return p


def resampler_params(p):
    pre = "model.resampler."
    return {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}


def inplace_splice_map(input_ids, tok: OmniTokens, num_query, image_of_slot=None):
    """omnilmm.py:219-258 as an index map: src[b][t] = t for text rows, -1 - (block*num_query + j) for the j-th row
    after an <im_start>. Blocks are consumed in batch order (cur_image_idx); raises like the reference when the
    matching <im_end> is missing."""
    nseq, L = input_ids.shape
    src = torch.arange(L, dtype=torch.int64).repeat(nseq, 1)
    slot = 0
    for b in range(nseq):
        row = input_ids[b]
        if int((row == tok.im_patch).sum()) == 0:
            continue
        starts = torch.where(row == tok.im_start)[0].tolist()
        if len(starts) != int((row == tok.im_end).sum()):
            raise ValueError("The number of image start tokens and image end tokens should be the same.")
        for s in starts:
            if s + num_query + 1 >= L or int(row[s + num_query + 1]) != tok.im_end:
                raise ValueError("The image end token should follow the image start token.")
            blk = slot if image_of_slot is None else int(image_of_slot[slot])
            src[b, s + 1:s + 1 + num_query] = -1 - (blk * num_query + torch.arange(num_query))
            slot += 1
    return src


def omnilmm_policy_logps(p, dec_cfg, res_cfg, tok, input_ids, labels, vision_tokens):
    """forward_DPO on the concatenated (win rows first) batch: vision_tokens [B, N, kv_dim] are the tower output
    WITHOUT prefix tokens; both halves of the batch see the same B images (trainers.py:190)."""
    nseq, L = input_ids.shape
    B = vision_tokens.shape[0]
    feats = R.resampler_forward(resampler_params(p), vision_tokens, res_cfg)           # [B, Q, H]
    Q = res_cfg.num_queries
    slots = torch.arange(nseq) % B if nseq == 2 * B else torch.arange(nseq)
    src = inplace_splice_map(input_ids, tok, Q, image_of_slot=slots)
    emb = p["model.embed_tokens.weight"][input_ids]                                    # [nseq, L, H]
    flat_feats = feats.reshape(B * Q, -1)
    is_img = src < 0
    embeds = torch.where(is_img.unsqueeze(-1), flat_feats[(-1 - src).clamp_min(0)], emb)
    logits = O.llama_logits(p, embeds, dec_cfg)
    per_tok, logp, avg = O.get_batch_logps(logits, labels)
    return dict(src=src, per_token_logps=per_tok, logp=logp, avg_logp=avg, logits=logits, image_features=feats)


def compute_weighted_logp(per_token_logp, labels, token_weight, use_average=False):
    """muffin/train/trainers.py:128-137 (the dpo_token_weighted branch of get_beta_and_logps, :246-261)."""
    weighted_mask = token_weight * (labels[:, 1:] != O.IGNORE_INDEX)
    logp = (per_token_logp * weighted_mask).sum(-1)
    return logp / weighted_mask.sum(-1) if use_average else logp


def omnilmm_dpo_step(p, dec_cfg, res_cfg, tok, batch, beta=0.1, use_average=False):
    """batch["token_weight"] [2B, L-1] present => the token-weighted loss (policy log-probs re-weighted per token;
    the reference log-probs in the batch are expected to be weighted the same way by the caller)."""
    out = omnilmm_policy_logps(p, dec_cfg, res_cfg, tok, batch["concatenated_input_ids"],
                               batch["concatenated_labels"], batch["vision_tokens"])
    if "token_weight" in batch:
        out["logp"] = compute_weighted_logp(out["per_token_logps"], batch["concatenated_labels"], batch["token_weight"],
                                            use_average)
    elif use_average:
        out["logp"] = out["avg_logp"]
    B = out["logp"].shape[0] // 2
    pw, pr = out["logp"][:B], out["logp"][B:]
    losses, cr, rr = O.dpo_loss(pw, pr, batch["ref_win_logp"], batch["ref_rej_logp"], beta)
    out.update(losses=losses, chosen_rewards=cr, rejected_rewards=rr, loss=losses.mean())
    return out


def synthetic_omni_batch(dec_cfg, res_cfg, tok, B, prompt_len, resp_len, seed, im_pos=4, ragged=True):
    """Pairs in the OmniLMM token layout: prompt = [text.. <im_start> <im_patch>*Q <im_end> text..], response
    supervised, right-padded with id 0 / label -100 (win rows first)."""
    g = torch.Generator().manual_seed(seed)
    Q = res_cfg.num_queries
    hi = min(tok.im_patch, tok.im_start, tok.im_end)
    rows_i, rows_l = [], []
    prompts = []
    for _ in range(B):
        pr = torch.randint(3, hi, (prompt_len,), generator=g)
        pr[0] = 1
        pr[im_pos] = tok.im_start
        pr[im_pos + 1:im_pos + 1 + Q] = tok.im_patch
        pr[im_pos + 1 + Q] = tok.im_end
        prompts.append(pr)
    for half in range(2):
        for i in range(B):
            n = resp_len - ((3 * i + 5 * half) % 7 if ragged else 0)
            resp = torch.randint(3, hi, (n,), generator=g)
            resp[-1] = 2
            rows_i.append(torch.cat([prompts[i], resp]))
            rows_l.append(torch.cat([torch.full((prompt_len,), O.IGNORE_INDEX), resp]))
    L = max(len(r) for r in rows_i)
    pad = lambda rows, v: torch.stack([torch.cat([r, torch.full((L - len(r),), v, dtype=torch.int64)]) for r in rows])
    return {"concatenated_input_ids": pad(rows_i, 0), "concatenated_labels": pad(rows_l, O.IGNORE_INDEX),
            "vision_tokens": torch.randn(B, res_cfg.kv_tokens, res_cfg.kv_dim, generator=g)}
