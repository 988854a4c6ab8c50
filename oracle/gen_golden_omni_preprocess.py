"""ORACLE TOOLING — fixture of the reference's OWN omni_preprocess / expand_image_token / encode_multimodal_preference_sample
(omnilmm/train/train_utils.py:50-151, muffin/train/train_utils.py:161-262) on a character-level chat tokenizer.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_omni_preprocess.py      (build container only)

The real tokenizer (sentencepiece + chat template of the OmniLMM checkpoint) is not available offline; the masking logic
only needs a tokenizer whose encoding of the two role templates is a sub-sequence of the encoded conversation, which a
character-level tokenizer guarantees. tests/test_omnilmm_data.py holds rlaifv_b200.omnilmm_data to this fixture.
"""
import os
import sys
from functools import partial

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CASES = [
    [{"from": "human", "value": "<image>\nwhat is shown ?"}, {"from": "gpt", "value": "a red bus ."}],
    [{"from": "human", "value": "describe <image> briefly"}, {"from": "gpt", "value": "two dogs"},
     {"from": "human", "value": "and the weather ?"}, {"from": "gpt", "value": "sunny , warm"}],
    [{"role": "user", "content": "no picture here"}, {"role": "assistant", "content": "indeed"},
     {"role": "user", "content": "a dangling question"}],
]


def main():
    from oracle import gen_golden as G
    from oracle.toy_tokenizer import CharChatTokenizer
    G.import_reference()                                     # puts /root/reference on sys.path, stubs matplotlib
    from omnilmm.train.train_utils import omni_preprocess as ref_pre
    from muffin.train.train_utils import encode_multimodal_preference_sample as ref_encode, expand_image_token as ref_expand
    tok = CharChatTokenizer()
    fx = {"n": np.int64(len(CASES))}
    import copy
    for i, conv in enumerate(CASES):
        for gen in (False, True):
            if gen and conv[-1].get("role", conv[-1].get("from")) in ("assistant", "gpt"):
                continue
            d = ref_pre([copy.deepcopy(conv)], tok, generation=gen)
            fx[f"c{i}_g{int(gen)}_ids"] = d["input_ids"][0].numpy()
            fx[f"c{i}_g{int(gen)}_labels"] = d["labels"][0].numpy()
    cfg = {"is_multimodal": True, "image_token_len": 4, "use_im_start_end": True,
           "image_processor": lambda im: torch.zeros(3, 2, 2)}
    src = {"question": {"from": "human", "value": "<image>\nhow many ?"}, "chosen": {"from": "gpt", "value": "three"},
           "rejected": {"from": "gpt", "value": "four apples"}, "image": "IMG", "ref_win_logp": -1.0, "ref_rej_logp": -2.0,
           "ref_win_avg_logp": -0.1, "ref_rej_avg_logp": -0.2, "ref_win_per_token_logp": [0.0] * 400,
           "ref_rej_per_token_logp": [0.0] * 400}
    rej, win = ref_encode(copy.deepcopy(src), tok, cfg, preprocess_func=ref_pre)
    fx["pair_win_ids"], fx["pair_win_labels"] = win["input_ids"].numpy(), win["labels"].numpy()
    fx["pair_rej_ids"], fx["pair_rej_labels"] = rej["input_ids"].numpy(), rej["labels"].numpy()
    exp = ref_expand([{"from": "human", "value": "a <image> b"}], cfg)
    fx["expanded"] = np.array(exp[0]["value"])
    out = os.path.join(REPO, "tests", "golden_host", "omni_preprocess.npz")
    np.savez_compressed(out, **fx)
    print("written", out, {k: v.shape for k, v in fx.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    main()
