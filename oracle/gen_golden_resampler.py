"""Pins oracle/resampler_oracle.py against the UNMODIFIED reference `Resampler`
(/root/reference/omnilmm/model/resampler.py, loaded by file path because `omnilmm/model/__init__` needs timm) and
writes tests/golden/resampler/resampler_*.npz. Build-container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_resampler.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import resampler_oracle as R  # noqa: E402

REF_FILE = "/root/reference/omnilmm/model/resampler.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_resampler", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sample(t, n=64):
    f = t.detach().flatten()
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy()


def main():
    ref = load_reference()
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden", "resampler")
    for name, B, seed in (("tiny_r", 3, 21), ("small_r", 2, 22)):
        cfg = R.R_CONFIGS[name]
        p = R.make_resampler_params(cfg, seed=seed)
        m = ref.Resampler(grid_size=cfg.grid_size, embed_dim=cfg.embed_dim, num_heads=cfg.num_heads, kv_dim=cfg.kv_dim)
        # the reference builds its own pos_embed: it must equal the restated table before we overwrite anything
        assert torch.allclose(m.pos_embed.data, p["pos_embed"], atol=1e-6)
        missing = m.load_state_dict(p, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        x, d_out = R.synthetic_vision_tokens(cfg, B, seed + 100)
        x_ref = x.clone().requires_grad_(True)
        y_ref = m(x_ref)
        y_ref.backward(d_out)
        po = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in p.items()}
        x_o = x.clone().requires_grad_(True)
        y_o = R.resampler_forward(po, x_o, cfg)
        y_o.backward(d_out)
        worst = float((y_o - y_ref).detach().abs().max() / y_ref.detach().abs().max())
        fx = {"cfg_name": name, "B": B, "seed": seed, "out_sample": sample(y_ref, 256),
              "out_norm": float(y_ref.norm()), "dx_sample": sample(x_ref.grad, 256), "dx_norm": float(x_ref.grad.norm())}
        if name == "tiny_r":
            fx["out_full"] = y_ref.detach().numpy()
        for k, prm in m.named_parameters():
            if k == "pos_embed":
                continue
            g_ref, g_o = prm.grad, po[k].grad
            e = float((g_o - g_ref).abs().max() / (g_ref.abs().max() + 1e-30))
            worst = max(worst, e)
            fx["gradsample:" + k] = sample(g_ref)
            fx["gradnorm:" + k] = float(g_ref.norm())
        e = float((x_o.grad - x_ref.grad).abs().max() / x_ref.grad.abs().max())
        worst = max(worst, e)
        print(f"{name}: restatement vs reference Resampler, worst relative error {worst:.2e}")
        assert worst < 2e-5, worst
        np.savez_compressed(os.path.join(out_dir, f"resampler_{name}.npz"), **fx)


if __name__ == "__main__":
    main()
