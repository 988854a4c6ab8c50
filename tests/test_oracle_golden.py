"""CPU tests: the oracle (restatement) against the golden fixtures produced by the unmodified
reference (oracle/gen_golden.py), plus known-answer properties of the path."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import llava_dpo_oracle as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def rel(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


_PARAMS = {}


def cfg_of(fx):
    return O.CONFIGS[str(fx["cfg_name"])] if "cfg_name" in fx.files else O.TINY


def params_for(scale, cfg=O.TINY):
    key = (scale, id(cfg))
    if key not in _PARAMS:
        _PARAMS[key] = O.make_params(cfg, seed=0, scale=scale)
    return _PARAMS[key]


@pytest.fixture(scope="module")
def params():
    return params_for(1.0)


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_outputs(path):
    fx = np.load(path)
    cfg = cfg_of(fx)
    params = params_for(float(fx["param_scale"]), cfg)
    assert abs(O.params_checksum(params) - float(fx["params_checksum"])) <= 1e-9 * float(fx["params_checksum"])
    p = {k: v.clone().requires_grad_(k.startswith(O.TRAINABLE_PREFIXES)) for k, v in params.items()}
    batch = dict(concatenated_input_ids=torch.from_numpy(fx["concatenated_input_ids"]),
                 concatenated_labels=torch.from_numpy(fx["concatenated_labels"]),
                 images=torch.from_numpy(fx["images"]),
                 ref_win_logp=torch.from_numpy(fx["ref_win_logp"]), ref_rej_logp=torch.from_numpy(fx["ref_rej_logp"]))
    out = O.dpo_step(p, cfg, batch, beta=float(fx["beta"]))
    # integer / copy work: bit exact
    assert torch.equal(out["labels"], torch.from_numpy(fx["spliced_labels"]))
    assert np.array_equal(out["inputs_embeds"].detach().double().sum(-1).numpy(), fx["spliced_embeds_rowsum"])
    # floating point: fp32 vs fp32
    assert rel(out["per_token_logps"].detach(), fx["per_token_logps"]) < 2e-5
    assert rel(out["policy_win_logp"].detach(), fx["policy_win_logp"]) < 2e-5
    assert rel(out["policy_rej_logp"].detach(), fx["policy_rej_logp"]) < 2e-5
    assert rel(out["losses"].detach(), fx["losses"]) < 2e-5
    assert rel(out["chosen_rewards"], fx["chosen_rewards"]) < 2e-5
    assert rel(out["rejected_rewards"], fx["rejected_rewards"]) < 2e-5
    assert abs(float(out["loss"]) - float(fx["loss"])) < 2e-5
    out["loss"].backward()
    for key in fx.files:
        if key.startswith("gradsample:"):
            name = key.split(":", 1)[1]
            g = p[name].grad.flatten()
            idx = torch.linspace(0, g.numel() - 1, 64).long()
            assert rel(g[idx], fx[key]) < 5e-5, name
            assert abs(float(g.double().norm()) - float(fx["gradnorm:" + name])) < 5e-5 * float(fx["gradnorm:" + name])
    # CLIP stays frozen (clip_encoder.py:46 no_grad)
    assert all(p[k].grad is None for k in p if "vision_tower" in k)


def test_padding_invariance_is_exact(params):
    """with attention_mask=None a padded, batched sequence gives bit-identical
    summed log-prob to the same sequence run alone (pads attend causally after all valid tokens)."""
    cfg = O.TINY
    batch = O.synthetic_pair_batch(cfg, 2, 20, 16, seed=3, image_pos=4, ragged=True)
    full = O.policy_logps(params, cfg, batch["concatenated_input_ids"], batch["concatenated_labels"], batch["images"])
    ids, labs = batch["concatenated_input_ids"], batch["concatenated_labels"]
    for row in range(ids.shape[0]):
        n = int((ids[row] != 0).sum())
        feats = O.clip_features(params, batch["images"][row % 2: row % 2 + 1], cfg)
        proj = O.mm_projector(params, feats)
        src, nl, T = O.splice_index_map(ids[row:row + 1, :n], labs[row:row + 1, :n], cfg.num_patches, cfg.max_len)
        emb = O.splice_embeds(params, ids[row:row + 1, :n], src, proj)
        logits = O.llama_logits(params, emb, cfg)
        _, lp, _ = O.get_batch_logps(logits, nl)
        assert abs(float(lp[0]) - float(full["logp"][row])) <= 2e-4 * abs(float(lp[0]))


def test_self_reference_loss_is_ln2(params):
    cfg = O.TINY
    batch = O.synthetic_pair_batch(cfg, 2, 20, 16, seed=9)
    out = O.policy_logps(params, cfg, batch["concatenated_input_ids"], batch["concatenated_labels"], batch["images"])
    B = 2
    losses, cr, rr = O.dpo_loss(out["logp"][:B], out["logp"][B:], out["logp"][:B].detach(), out["logp"][B:].detach(), 0.1)
    assert torch.allclose(losses, torch.full_like(losses, math.log(2.0)), atol=1e-6)
    assert float(cr.abs().max()) == 0.0 and float(rr.abs().max()) == 0.0


def test_splice_edge_cases():
    cfg = O.TINY
    P = cfg.num_patches
    ids = torch.tensor([[1, 5, -200, 7, 0, 0], [1, -200, 9, 10, 11, 2], [1, 4, 5, 6, 7, 2]])
    labs = torch.tensor([[-100, -100, -100, 7, -100, -100], [-100, -100, 9, 10, 11, 2], [-100, 4, 5, 6, 7, 2]])
    src, nl, T = O.splice_index_map(ids, labs, P, cfg.max_len)
    assert T == 6 - 1 + P
    assert src[0, 2].item() == -1 and src[0, 2 + P - 1].item() == -P          # block 0
    assert src[1, 1].item() == -1 - P                                          # block 1
    assert (src[2, :6] == torch.arange(6)).all() and (src[2, 6:] == -2 ** 31).all()   # no image: consumes block 2
    assert (nl[0, 2:2 + P] == -100).all() and nl[0, 2 + P].item() == 7
    # truncation to max_len
    src2, nl2, T2 = O.splice_index_map(ids, labs, P, 10)
    assert T2 == 10 and src2.shape == (3, 10)


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(64)
    g = torch.randn(64)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(64), torch.zeros(64)
    for step in range(1, 4):
        ref_p.grad = g.clone() * step
        opt.step()
        p, m, v = O.adamw_update(p, g * step, m, v, step, 1e-3)
        assert torch.allclose(p, ref_p.detach(), atol=1e-6)


def test_splice_edge_cases_match_reference_fixture():
    """Oracle splice on edge inputs (truncation, image-less sequence, image first/last, two images, very ragged) ==
    the unmodified reference's prepare_inputs_labels_for_multimodal (oracle/gen_golden_splice_edges.py)."""
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden_host", "splice_edge_cases.npz"))
    cfg = O.TINY
    params = O.make_params(cfg, seed=0)
    assert len(fx["names"]) == 6
    for name in [str(n) for n in fx["names"]]:
        ids, labels = torch.from_numpy(fx[name + ":ids"]), torch.from_numpy(fx[name + ":labels"])
        max_len, n_images = int(fx[name + ":max_len"]), int(fx[name + ":n_images"])
        g = torch.Generator().manual_seed(int(fx[name + ":image_seed"]))
        images = torch.randn(n_images, 3, cfg.image_size, cfg.image_size, generator=g)
        with torch.no_grad():
            feats = O.mm_projector(params, O.clip_features(params, images, cfg))
            src, new_labels, T = O.splice_index_map(ids, labels, cfg.num_patches, max_len)
            emb = O.splice_embeds(params, ids, src, feats)
        assert np.array_equal(new_labels.numpy(), fx[name + ":ref_labels"]), name            # bit-exact
        assert T <= max_len
        assert np.allclose(emb.double().sum(-1).numpy(), fx[name + ":ref_embeds_rowsum"], rtol=1e-6, atol=1e-6), name


LORA_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lora", "*.npz")))


@pytest.mark.parametrize("path", LORA_GOLDEN, ids=[os.path.basename(p) for p in LORA_GOLDEN])
def test_lora_oracle_matches_merged_weight_reference(path):
    """The oracle's LoRA branch (unmerged W x + s B(A x), what peft executes) against the unmodified reference model
    run with merged weights W' = W + s B A; adapter gradients via dA = s B^T dW', dB = s dW' A^T
    (oracle/gen_golden_lora.py)."""
    fx = np.load(path)
    cfg = O.TINY
    params = O.make_params(cfg, seed=0, scale=float(fx["param_scale"]))
    lora = O.make_lora_params(cfg, r=int(fx["r"]), seed=int(fx["lora_seed"]), b_std=float(fx["lora_b_std"]))
    assert abs(O.params_checksum(lora) - float(fx["lora_checksum"])) <= 1e-9 * float(fx["lora_checksum"])
    p = {k: v.clone().requires_grad_(k.startswith("model.mm_projector.")) for k, v in params.items()}
    p.update({k: v.clone().requires_grad_(True) for k, v in lora.items()})
    batch = dict(concatenated_input_ids=torch.from_numpy(fx["concatenated_input_ids"]),
                 concatenated_labels=torch.from_numpy(fx["concatenated_labels"]), images=torch.from_numpy(fx["images"]),
                 ref_win_logp=torch.from_numpy(fx["ref_win_logp"]), ref_rej_logp=torch.from_numpy(fx["ref_rej_logp"]))
    out = O.dpo_step(p, cfg, batch, beta=float(fx["beta"]))     # lora_scaling default 0.25 = alpha / r of the fixture
    assert float(fx["alpha"]) / int(fx["r"]) == 0.25
    assert torch.equal(out["labels"], torch.from_numpy(fx["spliced_labels"]))
    assert rel(out["per_token_logps"].detach(), fx["per_token_logps"]) < 5e-5
    assert rel(out["policy_win_logp"].detach(), fx["policy_win_logp"]) < 5e-5
    assert rel(out["losses"].detach(), fx["losses"]) < 5e-5
    out["loss"].backward()
    n = 0
    for key in fx.files:
        if key.startswith("gradsample:"):
            name = key.split(":", 1)[1]
            g = p[name].grad.flatten()
            idx = torch.linspace(0, g.numel() - 1, 64).long()
            assert rel(g[idx], fx[key]) < 1e-4, name
            assert abs(float(g.double().norm()) - float(fx["gradnorm:" + name])) < 1e-4 * float(fx["gradnorm:" + name])
            n += 1
    assert n == 2 * len(O.LORA_TARGETS) * cfg.num_layers + 4
