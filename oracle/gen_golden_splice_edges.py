"""ORACLE TOOLING — pins the oracle's image-token splice on EDGE inputs against the unmodified reference
`prepare_inputs_labels_for_multimodal` (llava/model/llava_arch.py:150-330, attention_mask=None path):
truncation at tokenizer_model_max_length (:280-283), a sequence without an image token (:239-246: consumes a feature
block, emits no rows), image token first / last, responses of very different lengths, and two image tokens in one
sequence. Writes tests/golden_host/splice_edge_cases.npz (inputs + the reference's spliced labels and the row sums of
its spliced embeddings). Build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_splice_edges.py
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
from oracle import llava_dpo_oracle as O       # noqa: E402
import gen_golden as G                         # noqa: E402


def cases(cfg):
    g = torch.Generator().manual_seed(99)
    P = cfg.num_patches

    def seq(n, img_at, n_lab):
        ids = torch.randint(3, cfg.vocab_size, (n,), generator=g)
        ids[0] = 1
        for j in img_at:
            ids[j] = O.IMAGE_TOKEN_INDEX
        lab = torch.full((n,), O.IGNORE_INDEX)
        lab[n - n_lab:] = ids[n - n_lab:]
        return ids, lab

    def pad(rows, L=None):
        L = L or max(len(i) for i, _ in rows)
        ids = torch.stack([torch.cat([i, torch.zeros(L - len(i), dtype=torch.int64)]) for i, _ in rows])
        lab = torch.stack([torch.cat([l, torch.full((L - len(l),), O.IGNORE_INDEX)]) for _, l in rows])
        return ids, lab

    out = {}
    out["truncate_max_len_20"] = (*pad([seq(30, [4], 12), seq(18, [2], 6)]), 20, 2)           # both sides of the cut
    out["truncate_inside_image"] = (*pad([seq(12, [9], 2), seq(12, [1], 5)]), 9 + P - 1, 2)   # cut falls in image rows
    out["no_image_sequence"] = (*pad([seq(14, [3], 5), seq(16, [], 7), seq(10, [6], 3)]), 2048, 3)
    out["image_first_and_last"] = (*pad([seq(9, [0], 4), seq(11, [10], 0)]), 2048, 2)
    out["two_images_one_sequence"] = (*pad([seq(15, [2, 8], 4), seq(13, [5], 6)]), 2048, 3)
    out["very_ragged"] = (*pad([seq(40, [7], 30), seq(6, [1], 2), seq(21, [20], 0)]), 2048, 3)
    return out


def main():
    R = G.import_reference()
    cfg = O.TINY
    params = O.make_params(cfg, seed=0)
    model = G.build_reference_model(R, cfg, params)
    model.eval()
    fx = {}
    names = []
    for name, (ids, labels, max_len, n_images) in cases(cfg).items():
        model.config.tokenizer_model_max_length = max_len
        g = torch.Generator().manual_seed(len(name))
        images = torch.randn(n_images, 3, cfg.image_size, cfg.image_size, generator=g)
        with torch.no_grad():
            _, _, _, _, ref_embeds, ref_labels = model.prepare_inputs_labels_for_multimodal(
                input_ids=ids.clone(), position_ids=None, attention_mask=None, past_key_values=None,
                labels=labels.clone(), images=images)
            # oracle on the same inputs
            feats = O.mm_projector(params, O.clip_features(params, images, cfg))
            src, new_labels, T = O.splice_index_map(ids, labels, cfg.num_patches, max_len)
            emb = O.splice_embeds(params, ids, src, feats)
        assert torch.equal(new_labels, ref_labels), name
        assert torch.equal(emb, ref_embeds), name
        print(f"{name}: T={T}, labels and spliced rows bit-exact vs the reference")
        names.append(name)
        fx[name + ":ids"], fx[name + ":labels"] = ids.numpy(), labels.numpy()
        fx[name + ":max_len"], fx[name + ":n_images"] = np.int64(max_len), np.int64(n_images)
        fx[name + ":image_seed"] = np.int64(len(name))
        fx[name + ":ref_labels"] = ref_labels.numpy()
        fx[name + ":ref_embeds_rowsum"] = ref_embeds.double().sum(-1).numpy()
    fx["names"] = np.array(names)
    np.savez_compressed(os.path.join(REPO, "tests", "golden_host", "splice_edge_cases.npz"), **fx)


if __name__ == "__main__":
    main()
