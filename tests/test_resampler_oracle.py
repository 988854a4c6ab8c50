"""CPU: the resampler restatement (oracle/resampler_oracle.py) against the fixtures written from the UNMODIFIED
reference `Resampler` class (oracle/gen_golden_resampler.py -> tests/golden/resampler/*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import resampler_oracle as R

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "resampler", "*.npz")))


def sample(t, n=64):
    f = t.detach().flatten()
    return f[torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()].numpy()


def test_fixtures_present():
    assert len(GOLDEN) == 2


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_resampler_fixture(path):
    fx = np.load(path)
    cfg = R.R_CONFIGS[str(fx["cfg_name"])]
    p = R.make_resampler_params(cfg, seed=int(fx["seed"]))
    p = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in p.items()}
    x, d_out = R.synthetic_vision_tokens(cfg, int(fx["B"]), int(fx["seed"]) + 100)
    x.requires_grad_(True)
    y = R.resampler_forward(p, x, cfg)
    y.backward(d_out)
    tol = 2e-5
    assert np.abs(sample(y, 256) - fx["out_sample"]).max() <= tol * np.abs(fx["out_sample"]).max()
    assert abs(float(y.norm()) - float(fx["out_norm"])) <= tol * float(fx["out_norm"])
    if "out_full" in fx.files:
        assert np.abs(y.detach().numpy() - fx["out_full"]).max() <= tol * np.abs(fx["out_full"]).max()
    assert np.abs(sample(x.grad, 256) - fx["dx_sample"]).max() <= tol * np.abs(fx["dx_sample"]).max()
    for key in fx.files:
        if key.startswith("gradsample:"):
            name = key.split(":", 1)[1]
            ref = fx[key]
            assert np.abs(sample(p[name].grad) - ref).max() <= 5e-5 * (np.abs(ref).max() + 1e-12), name
            gn = float(fx["gradnorm:" + name])
            assert abs(float(p[name].grad.norm()) - gn) <= 5e-5 * gn, name


def test_position_table_properties():
    t = R.sincos_2d(64, 4)
    assert t.shape == (16, 64) and float(t.abs().max()) <= 1.0
    # row 0 = position (0, 0): sin parts 0, cos parts 1
    assert torch.allclose(t[0, :16], torch.zeros(16)) and torch.allclose(t[0, 16:32], torch.ones(16))
    assert R.abs_pos(t, 16) is t and R.abs_pos(t, 144).shape == (144, 64)
