"""Host-side sample encoding of the OmniLMM-12B (RLAIF-V-12B) DPO path — the caller side of BASELINE config (d).

Mirrors, with the same names and return layouts:
  omni_preprocess            omnilmm/train/train_utils.py:50-151   chat-template tokenisation + label masking: only the
                                                                  assistant turns are supervised
  expand_image_token         muffin/train/train_utils.py:161-175   "<image>" -> <im_start> <im_patch>*num_query <im_end>
  encode_omni_preference_sample  = encode_multimodal_preference_sample (muffin/train/train_utils.py:198-262) with
                                   preprocess_func=omni_preprocess and use_im_start_end=True, the combination the
                                   OmniLMM policy's in-place splice expects (omnilmm/model/omnilmm.py:219-258)
Pure tokenizer / list work (DataLoader workers); nothing here touches the GPU.
"""
import copy
import warnings

import numpy as np
import torch

IGNORE_INDEX = -100
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"
SYSTEM_PROMPT = ("You are an artificial intelligence assistant, which gives helpful, detailed, and polite answers to the "
                 "human's questions.")
RESPONSE_TEMPLATE = "\n<|assistant|>\n"
INSTRUCTION_TEMPLATE = "\n<|user|>\n"


def _tokenize_one(text, tokenizer):
    enc = tokenizer(text, return_tensors="pt", padding="longest", max_length=tokenizer.model_max_length, truncation=True)
    return enc.input_ids[0]


def _find_all(haystack, needle):
    """Start indices where the id list `needle` occurs in the 1-D tensor `haystack`."""
    hits = []
    n = len(needle)
    for i in np.where(haystack.numpy() == needle[0])[0]:
        if haystack[i:i + n].tolist() == needle:
            hits.append(int(i))
    return hits


def omni_preprocess(sources, tokenizer, generation=False):
    """sources: list of conversations (lists of {'from'|'role', 'value'|'content'} turns). Returns
    dict(input_ids=[LongTensor], labels=[LongTensor]); labels are -100 except inside assistant turns."""
    resp_ids = tokenizer.encode(RESPONSE_TEMPLATE, add_special_tokens=False)
    inst_ids = tokenizer.encode(INSTRUCTION_TEMPLATE, add_special_tokens=False)
    all_ids, all_labels = [], []
    for conv in sources:
        turns, prev = [], None
        for t in conv:
            role = t["from"] if "from" in t else t["role"]
            role = {"human": "user", "gpt": "assistant"}.get(role, role)
            assert role in ("user", "assistant") and role != prev, f"role={role}, prev_role={prev}"
            prev = role
            turns.append({"role": role, "content": t["value"] if "value" in t else t["content"]})
        if turns[0]["role"] != "system":
            turns.insert(0, {"role": "system", "content": SYSTEM_PROMPT})
        text = tokenizer.apply_chat_template(turns, tokenize=False, add_generation_prompt=generation)
        if not generation:
            text = text.strip()
        ids = _tokenize_one(text, tokenizer)
        labels = copy.deepcopy(ids)
        resp_starts = [i + len(resp_ids) for i in _find_all(labels, resp_ids)]       # first token AFTER the template
        user_starts = _find_all(labels, inst_ids)
        if not resp_starts or not user_starts:
            warnings.warn("omni_preprocess: response / instruction template not found; the instance is ignored in the "
                          "loss (consider a larger max length)")
            labels[:] = IGNORE_INDEX
        for k, (u, r) in enumerate(zip(user_starts, resp_starts)):
            if k == 0:
                labels[:r] = IGNORE_INDEX            # system prompt + first question + the template itself
            else:
                labels[u:r] = IGNORE_INDEX           # a later question
        if len(resp_starts) < len(user_starts):
            labels[user_starts[-1]:] = IGNORE_INDEX  # trailing question without an answer
        all_ids.append(ids)
        all_labels.append(labels)
    return dict(input_ids=all_ids, labels=all_labels)


def expand_image_token(source, multimodal_cfg):
    if not multimodal_cfg["is_multimodal"] or multimodal_cfg.get("keep_image_tag", False):
        return source
    rep = DEFAULT_IMAGE_PATCH_TOKEN * multimodal_cfg["image_token_len"]
    if multimodal_cfg["use_im_start_end"]:
        rep = DEFAULT_IM_START_TOKEN + rep + DEFAULT_IM_END_TOKEN
    for sentence in source:
        sentence["value"] = sentence["value"].replace(DEFAULT_IMAGE_TOKEN, rep)
    return source


def encode_omni_preference_sample(source, tokenizer, multimodal_cfg):
    """-> (rej_dict, win_dict) with input_ids, labels, image and the cached reference log-probs — the tuple the
    preference collator consumes (muffin/train/train_muffin.py:43-112)."""
    if isinstance(source["chosen"], list):
        win_conv, rej_conv = source["chosen"], source["rejected"]
    else:
        win_conv = copy.deepcopy([source["question"], source["chosen"]])
        rej_conv = copy.deepcopy([source["question"], source["rejected"]])
    image = None
    if "image" in source:
        image = multimodal_cfg["image_processor"](source["image"])
        win_conv = expand_image_token(win_conv, multimodal_cfg)
        rej_conv = expand_image_token(rej_conv, multimodal_cfg)
    out = []
    for conv in (rej_conv, win_conv):
        enc = omni_preprocess([conv], tokenizer)
        out.append({"input_ids": enc["input_ids"][0], "labels": enc["labels"][0]})
    rej, win = out
    if image is not None:
        rej["image"] = win["image"] = image
    elif multimodal_cfg["is_multimodal"]:
        cs = multimodal_cfg["image_processor"].crop_size
        rej["image"] = win["image"] = torch.zeros(3, cs["height"], cs["width"])
    if "ref_win_logp" in source:
        for k in ("logp", "avg_logp", "per_token_logp"):
            rej[f"ref_rej_{k}"] = source[f"ref_rej_{k}"]
            win[f"ref_win_{k}"] = source[f"ref_win_{k}"]
    return rej, win
