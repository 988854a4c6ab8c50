"""CPU: the OmniLMM restatement (oracle/omnilmm_oracle.py) against fixtures written from the UNMODIFIED reference
OmniLMMForCausalLM + forward_DPO + dpo_loss (oracle/gen_golden_omnilmm.py -> tests/golden/omnilmm/*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import omnilmm_oracle as OM

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "omnilmm", "*.npz")))


def sample(t, n=64):
    f = t.detach().flatten()
    return f[torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()].numpy()


def test_fixtures_present():
    assert len(GOLDEN) == 3 and any("weighted" in p for p in GOLDEN)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_omnilmm_fixture(path):
    fx = np.load(path)
    dec, res, tok = OM.TINY_OMNI_DEC, OM.TINY_OMNI_RES, OM.TINY_OMNI_TOK
    B, seed = int(fx["B"]), int(fx["seed"])
    p = {k: v.clone().requires_grad_("pos_embed" not in k) for k, v in OM.make_omnilmm_params(dec, res, seed).items()}
    batch = OM.synthetic_omni_batch(dec, res, tok, B, 28, 20, seed=seed + 7, ragged=bool(fx["ragged"]))
    batch["vision_tokens"].requires_grad_(True)
    batch["ref_win_logp"] = torch.from_numpy(fx["ref_win_logp"])
    batch["ref_rej_logp"] = torch.from_numpy(fx["ref_rej_logp"])
    if "token_weight" in fx.files:                       # --dpo_token_weighted case (trainers.py:246-261)
        batch["token_weight"] = torch.from_numpy(fx["token_weight"])
    o = OM.omnilmm_dpo_step(p, dec, res, tok, batch)
    o["loss"].backward()
    close = lambda a, b, tol: np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * (np.abs(np.asarray(b)).max() + 1e-30)
    assert close(o["logp"].detach().numpy(), fx["logp"], 1e-5)
    assert close(o["per_token_logps"].detach().numpy(), fx["per_token_logps"], 1e-5)
    assert close(o["losses"].detach().numpy(), fx["losses"], 1e-5)
    assert close(o["chosen_rewards"].detach().numpy(), fx["chosen_rewards"], 1e-4)
    assert close(sample(batch["vision_tokens"].grad, 256), fx["dvision_sample"], 1e-4)
    for key in fx.files:
        if key.startswith("gradsample:"):
            name = key.split(":", 1)[1]
            assert close(sample(p[name].grad), fx[key], 1e-4), name


def test_inplace_splice_map_semantics():
    tok, Q = OM.TINY_OMNI_TOK, 4
    ids = torch.tensor([[1, 7, tok.im_start, tok.im_patch, tok.im_patch, tok.im_patch, tok.im_patch, tok.im_end, 9, 0],
                        [1, 5, 6, 7, 8, 9, 10, 11, 2, 0],                      # text-only: consumes no image
                        [tok.im_start, tok.im_patch, tok.im_patch, tok.im_patch, tok.im_patch, tok.im_end, 4, 5, 6, 2]])
    src = OM.inplace_splice_map(ids, tok, Q)
    assert src[0].tolist() == [0, 1, 2, -1, -2, -3, -4, 7, 8, 9]
    assert src[1].tolist() == list(range(10))
    assert src[2].tolist() == [0, -5, -6, -7, -8, 5, 6, 7, 8, 9]              # second image block
    bad = ids.clone()
    bad[0, 7] = 9                                                              # <im_end> missing
    with pytest.raises(ValueError):
        OM.inplace_splice_map(bad, tok, Q)


def test_host_compute_weighted_logp_matches_reference_output():
    """rlaifv_b200.trainers.compute_weighted_logp (host side of --dpo_token_weighted) on the reference's own per-token
    log-probs reproduces the reference's compute_weighted_logp output stored in the fixture."""
    from rlaifv_b200.trainers import compute_weighted_logp
    fx = np.load([p for p in GOLDEN if "weighted" in p][0])
    dec, res, tok = OM.TINY_OMNI_DEC, OM.TINY_OMNI_RES, OM.TINY_OMNI_TOK
    batch = OM.synthetic_omni_batch(dec, res, tok, int(fx["B"]), 28, 20, seed=int(fx["seed"]) + 7, ragged=True)
    pt, tw = torch.from_numpy(fx["per_token_logps"]), torch.from_numpy(fx["token_weight"])
    got = compute_weighted_logp(pt, batch["concatenated_labels"], tw, False)
    assert torch.allclose(got, torch.from_numpy(fx["logp"]), rtol=1e-6, atol=1e-4)
    avg = compute_weighted_logp(pt, batch["concatenated_labels"], tw, True)
    wm = tw * (batch["concatenated_labels"][:, 1:] != -100)
    assert torch.allclose(avg, got / wm.sum(-1), rtol=1e-6)
