"""Preference collator — host-side mirror of the reference's batch builder, same names, keys,
ordering and padding rules so the reference's dataset objects plug in unchanged:

  SFT_collator_fn            muffin/train/train_utils.py:55-96
  concate_pad                muffin/eval/muffin_inference_logp.py:180-185
  preference_collator_fn     muffin/eval/muffin_inference_logp.py:187-208   (win rows first, then rej)
  DataCollatorForDPODataset  muffin/train/train_muffin.py:37-112            (20 keys)
  get_diff_ids               utils/diff_lib.py:114-178                     (difflib token diff)

Pure host code (tokens, python lists, small CPU tensors): nothing here is a kernel target.
"""
import difflib
from dataclasses import dataclass
from typing import Any, Dict, Sequence

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence

IGNORE_INDEX = -100


def SFT_collator_fn(instances, pad_token_id):
    ids = [ins["input_ids"] for ins in instances]
    labs = [ins["labels"] for ins in instances]
    input_ids = pad_sequence(ids, batch_first=True, padding_value=pad_token_id)
    labels = pad_sequence(labs, batch_first=True, padding_value=IGNORE_INDEX)
    batch = {"input_ids": input_ids, "labels": labels, "attention_mask": input_ids.ne(pad_token_id)}
    imgs = [ins["image"] for ins in instances if "image" in ins]
    if not imgs:
        batch["images"] = []
    elif imgs[0].ndim == 4:                       # several images per sample: keep the list
        batch["images"] = imgs
    elif all(im is not None and im.shape == imgs[0].shape for im in imgs):
        batch["images"] = torch.stack([torch.from_numpy(im) if isinstance(im, np.ndarray) else im for im in imgs])
    else:
        batch["images"] = imgs
    if "context_ids" in instances[0]:             # MiniCPM-V branch of the reference
        batch["image_bounds"] = [ins["image_bounds"] for ins in instances]
        batch["context_ids"] = pad_sequence([ins["context_ids"] for ins in instances], batch_first=True,
                                            padding_value=0)
    return batch


def concate_pad(tensorA, tensorB, padding_value):
    return pad_sequence(list(tensorA) + list(tensorB), batch_first=True, padding_value=padding_value)


def preference_collator_fn(instances, pad_token_id):
    rej, win = zip(*instances)                    # instances are (rej_dict, win_dict) tuples
    rb, wb = SFT_collator_fn(rej, pad_token_id), SFT_collator_fn(win, pad_token_id)
    cat_ids = concate_pad(wb["input_ids"], rb["input_ids"], pad_token_id)
    return {
        "concatenated_input_ids": cat_ids,
        "concatenated_labels": concate_pad(wb["labels"], rb["labels"], IGNORE_INDEX),
        "concatenated_attention_mask": cat_ids.ne(pad_token_id),
        "win_input_ids": wb["input_ids"], "rej_input_ids": rb["input_ids"],
        "win_labels": wb["labels"], "rej_labels": rb["labels"],
        "win_attention_mask": wb["attention_mask"], "rej_attention_mask": rb["attention_mask"],
        "images": wb["images"],
    }


# ---- token-level diff used for the per-token DPO weights -----------------------------------------
def _kept_matches(a_seq, b_seq, min_match_size):
    blocks = difflib.SequenceMatcher(None, a_seq, b_seq).get_matching_blocks()
    kept = [m for m in blocks[:-1] if m.size >= min_match_size] + [blocks[-1]]   # sentinel always kept
    return [(m.a, m.a + m.size) for m in kept], [(m.b, m.b + m.size) for m in kept]


def _gaps(matches, length):
    """Span i = the stretch between kept match i-1 and kept match i (span 0 starts at 0). The last
    kept match is difflib's (len_a, len_b, 0) sentinel, so the final stretch is covered too."""
    gaps, start = [], 0
    for s, e in matches:
        gaps.append((start, s))
        start = e
    return gaps


def get_diff_ids(a_seq, b_seq, min_match_size=3):
    """Positions of a_seq / b_seq inside *modified* spans: gaps between matching blocks that are
    non-empty on both sides."""
    am, bm = _kept_matches(a_seq, b_seq, min_match_size)
    a_ids, b_ids = set(), set()
    for (a0, a1), (b0, b1) in zip(_gaps(am, len(a_seq)), _gaps(bm, len(b_seq))):
        if a0 != a1 and b0 != b1:
            a_ids.update(range(a0, a1))
            b_ids.update(range(b0, b1))
    return sorted(a_ids), sorted(b_ids)


@dataclass
class DataCollatorForDPODataset:
    tokenizer: Any
    beta: float
    mod_token_weight: float
    # Extension (off = the reference's 20 keys, bit-exact): also emit the UN-truncated reference per-token log-probs.
    # For LLaVA-1.5 the cached lists are in spliced positions (length n - 1 + 575) and the reference's cut at L - 1
    # (train_muffin.py:78-81) drops their tail — one of the two reasons it refuses --dpo_token_weighted for LLaVA.
    keep_spliced_per_token: bool = False

    def __call__(self, instances: Sequence[Dict]) -> Dict[str, torch.Tensor]:
        batch = preference_collator_fn(instances, self.tokenizer.pad_token_id)
        rej, win = zip(*instances)
        batch["beta"] = self.beta
        for kind, group in (("win", win), ("rej", rej)):
            batch[f"ref_{kind}_logp"] = torch.as_tensor([x[f"ref_{kind}_logp"] for x in group])
            batch[f"ref_{kind}_avg_logp"] = torch.as_tensor([x[f"ref_{kind}_avg_logp"] for x in group])
        per_tok = {}
        for kind, group in (("win", win), ("rej", rej)):
            seqs = [torch.as_tensor(x[f"ref_{kind}_per_token_logp"]) for x in group]
            padded = pad_sequence(seqs, batch_first=True, padding_value=0)
            need = batch[f"{kind}_input_ids"].size(1) - 1     # logits of the last token are unused
            assert padded.size(1) >= need, f"{padded.size(1)} >= {need}"
            batch[f"ref_{kind}_per_token_logp"] = padded[:, :need]
            if self.keep_spliced_per_token:
                batch[f"ref_{kind}_per_token_logp_spliced"] = padded
            per_tok[kind] = torch.ones_like(batch[f"ref_{kind}_per_token_logp"])
        for i, (w, r) in enumerate(zip(batch["win_input_ids"], batch["rej_input_ids"])):
            r_mod, w_mod = get_diff_ids(r[1:].tolist(), w[1:].tolist(), min_match_size=3)
            per_tok["win"][i][w_mod] = self.mod_token_weight
            per_tok["rej"][i][r_mod] = self.mod_token_weight
        batch["win_token_weight"] = per_tok["win"]
        batch["rej_token_weight"] = per_tok["rej"]
        batch["concatenated_token_weight"] = concate_pad(per_tok["win"], per_tok["rej"], 0)
        for ins in list(win) + list(rej):
            assert len(ins["input_ids"]) == len(ins["labels"])
        if torch.isnan(batch["win_token_weight"]).any() or torch.isnan(batch["rej_token_weight"]).any():
            raise FloatingPointError("NaN in token weights")
        return batch
