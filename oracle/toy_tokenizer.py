"""TEST INFRASTRUCTURE: whitespace tokenizer with BOS/EOS, used to exercise the llava_v1 template and
label-masking logic of both the reference (in oracle/gen_golden.py) and this repo without the real
sentencepiece model (not available offline)."""
from types import SimpleNamespace


class ToyTokenizer:
    bos_token_id, pad_token_id, model_max_length, legacy = 1, 0, 2048, True

    def __init__(self):
        self.vocab = {}

    def __call__(self, text):
        ids = [1]
        for w in text.replace("</s>", " </s> ").split():
            ids.append(2 if w == "</s>" else self.vocab.setdefault(w, 3 + len(self.vocab)))
        return SimpleNamespace(input_ids=ids)


class CharChatTokenizer:
    """TEST INFRASTRUCTURE: character-level tokenizer with a zephyr-style chat template, for the OmniLMM sample
    encoding (omni_preprocess). Every character is one id (3 + code point), so the encoding of a role template is a
    sub-sequence of the encoded conversation — the only property the masking logic relies on. The three image
    placeholder strings are single ids like the added special tokens of a real OmniLMM tokenizer."""
    pad_token_id, bos_token_id, eos_token_id, model_max_length = 0, 1, 2, 4096
    SPECIALS = {"<im_patch>": 500000, "<im_start>": 500001, "<im_end>": 500002, "</s>": 2}

    def apply_chat_template(self, messages, tokenize=False, add_generation_prompt=False):
        assert not tokenize
        text = "".join("<|%s|>\n%s</s>\n" % (m["role"], m["content"]) for m in messages)
        return text + ("<|assistant|>\n" if add_generation_prompt else "")

    def _ids(self, text):
        ids, i = [], 0
        while i < len(text):
            for sp, sid in self.SPECIALS.items():
                if text.startswith(sp, i):
                    ids.append(sid)
                    i += len(sp)
                    break
            else:
                ids.append(3 + ord(text[i]))
                i += 1
        return ids

    def encode(self, text, add_special_tokens=True):
        return ([self.bos_token_id] if add_special_tokens else []) + self._ids(text)

    def __call__(self, text, return_tensors=None, padding=None, max_length=None, truncation=False):
        import torch
        ids = self.encode(text)
        if truncation and max_length:
            ids = ids[:max_length]
        return SimpleNamespace(input_ids=torch.tensor([ids], dtype=torch.long) if return_tensors == "pt" else ids)

    def decode(self, ids):
        return "".join(chr(int(i) - 3) if 3 <= int(i) < 500000 else "?" for i in ids)
