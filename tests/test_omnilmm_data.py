"""CPU: the OmniLMM host-side sample encoding (rlaifv_b200.omnilmm_data) against a fixture produced by the reference's
own omni_preprocess / expand_image_token / encode_multimodal_preference_sample (oracle/gen_golden_omni_preprocess.py)
— token ids and label masks bit-exact."""
import copy
import os

import numpy as np
import torch

from oracle.gen_golden_omni_preprocess import CASES
from oracle.toy_tokenizer import CharChatTokenizer
from rlaifv_b200 import omnilmm_data as D

FX = np.load(os.path.join(os.path.dirname(__file__), "golden_host", "omni_preprocess.npz"))


def test_omni_preprocess_ids_and_label_masks_bit_exact():
    tok = CharChatTokenizer()
    n = 0
    for i, conv in enumerate(CASES):
        for gen in (False, True):
            key = f"c{i}_g{int(gen)}_ids"
            if key not in FX.files:
                continue
            d = D.omni_preprocess([copy.deepcopy(conv)], tok, generation=gen)
            assert np.array_equal(d["input_ids"][0].numpy(), FX[key]), key
            assert np.array_equal(d["labels"][0].numpy(), FX[f"c{i}_g{int(gen)}_labels"]), key
            n += 1
    assert n == 4
    # only assistant text is supervised: the two-turn case keeps both answers, masks both questions
    lab = torch.from_numpy(FX["c1_g0_labels"])
    text = tok.decode(lab[lab != -100])
    assert "two dogs" in text and "sunny , warm" in text and "weather" not in text and "describe" not in text
    # the dangling final question contributes nothing
    lab = torch.from_numpy(FX["c2_g0_labels"])
    assert "dangling" not in tok.decode(lab[lab != -100])


def test_image_expansion_and_pair_encoding_match_reference():
    tok = CharChatTokenizer()
    cfg = {"is_multimodal": True, "image_token_len": 4, "use_im_start_end": True,
           "image_processor": lambda im: torch.zeros(3, 2, 2)}
    exp = D.expand_image_token([{"from": "human", "value": "a <image> b"}], cfg)
    assert exp[0]["value"] == str(FX["expanded"]) == "a <im_start><im_patch><im_patch><im_patch><im_patch><im_end> b"
    src = {"question": {"from": "human", "value": "<image>\nhow many ?"}, "chosen": {"from": "gpt", "value": "three"},
           "rejected": {"from": "gpt", "value": "four apples"}, "image": "IMG", "ref_win_logp": -1.0, "ref_rej_logp": -2.0,
           "ref_win_avg_logp": -0.1, "ref_rej_avg_logp": -0.2, "ref_win_per_token_logp": [0.0] * 400,
           "ref_rej_per_token_logp": [0.0] * 400}
    rej, win = D.encode_omni_preference_sample(copy.deepcopy(src), tok, cfg)
    assert np.array_equal(win["input_ids"].numpy(), FX["pair_win_ids"]) and np.array_equal(win["labels"].numpy(), FX["pair_win_labels"])
    assert np.array_equal(rej["input_ids"].numpy(), FX["pair_rej_ids"]) and np.array_equal(rej["labels"].numpy(), FX["pair_rej_labels"])
    ids = win["input_ids"].tolist()
    s = ids.index(tok.SPECIALS["<im_start>"])
    assert ids[s + 1:s + 5] == [tok.SPECIALS["<im_patch>"]] * 4 and ids[s + 5] == tok.SPECIALS["<im_end>"]   # in-place splice layout
    assert win["ref_win_logp"] == -1.0 and rej["ref_rej_logp"] == -2.0 and win["image"].shape == (3, 2, 2)
