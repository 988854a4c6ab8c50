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
