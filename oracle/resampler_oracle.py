"""TEST INFRASTRUCTURE ONLY — CPU restatement of the OmniLMM perceiver resampler (BASELINE config d,
SURVEY.md §8 a13). Only tests/ may import this; the product path (rlaif-v_b200/resampler.py) never does.

Follows /root/reference/omnilmm/model/resampler.py:
  get_2d_sincos_pos_embed  :42-87      frozen 2-D sin-cos position table of the grid_size² learned queries
  get_abs_pos              :23-39      bicubic resize of that table to the number of vision tokens
  Resampler.__init__       :101-135    query, kv_proj (no bias), nn.MultiheadAttention, ln_q / ln_kv / ln_post
                                       (LayerNorm eps 1e-6), proj
  Resampler.forward        :149-168    x = ln_kv(kv_proj(x)); q = ln_q(query);
                                       out = MHA(q + pos_q, x + pos_kv, x); out = ln_post(out) @ proj
nn.MultiheadAttention is third-party (torch): restated from its documented arithmetic (packed in_proj rows = q|k|v,
scaled dot-product softmax(q kᵀ/√d) v per head, out_proj), and PINNED: oracle/gen_golden_resampler.py runs the
unmodified reference `Resampler` class (loaded straight from the file above) on the same seeded weights / inputs
and writes tests/golden/resampler/resampler_*.npz (outputs, input gradient, every parameter gradient).
"""
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class ResamplerConfig:
    grid_size: int = 8          # 64 queries (omnilmm/model/omnilmm.py:46-51: sqrt(num_query))
    embed_dim: int = 4096       # LLM hidden size
    num_heads: int = 32         # embed_dim // 128
    kv_dim: int = 1792          # EVA-02-E width
    kv_tokens: int = 1024       # 448 px / 14 = 32 x 32 patches
    eps: float = 1e-6

    @property
    def num_queries(self):
        return self.grid_size ** 2

    @property
    def head_dim(self):
        return self.embed_dim // self.num_heads


TINY_R = ResamplerConfig(grid_size=4, embed_dim=256, num_heads=2, kv_dim=192, kv_tokens=144)
SMALL_R = ResamplerConfig(grid_size=8, embed_dim=512, num_heads=4, kv_dim=320, kv_tokens=1024)
R_CONFIGS = {"tiny_r": TINY_R, "small_r": SMALL_R}

PARAM_SHAPES = lambda c: {                                    # names = Resampler.named_parameters()
    "query": (c.num_queries, c.embed_dim),
    "proj": (c.embed_dim, c.embed_dim),
    "kv_proj.weight": (c.embed_dim, c.kv_dim),
    "attn.in_proj_weight": (3 * c.embed_dim, c.embed_dim),
    "attn.in_proj_bias": (3 * c.embed_dim,),
    "attn.out_proj.weight": (c.embed_dim, c.embed_dim),
    "attn.out_proj.bias": (c.embed_dim,),
    "ln_q.weight": (c.embed_dim,), "ln_q.bias": (c.embed_dim,),
    "ln_kv.weight": (c.embed_dim,), "ln_kv.bias": (c.embed_dim,),
    "ln_post.weight": (c.embed_dim,), "ln_post.bias": (c.embed_dim,),
}


def sincos_1d(embed_dim, pos):
    """resampler.py:72-87: [M] positions -> [M, embed_dim] = [sin | cos] of pos * 10000^(-i/(D/2))."""
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim, grid_size):
    """resampler.py:42-69: first half of the channels encodes the w index, second half the h index (meshgrid with w
    first)."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_size, grid_size)
    emb_h = sincos_1d(embed_dim // 2, grid[0])
    emb_w = sincos_1d(embed_dim // 2, grid[1])
    return torch.from_numpy(np.concatenate([emb_h, emb_w], axis=1)).float()


def abs_pos(pos_embed, tgt_len):
    """resampler.py:23-39: bicubic (align_corners=False) resize of the [L,C] table to tgt_len tokens."""
    src = int(math.sqrt(pos_embed.shape[0]))
    tgt = int(math.sqrt(tgt_len))
    if src == tgt:
        return pos_embed
    dt = pos_embed.dtype
    t = pos_embed.float().reshape(1, src, src, -1).permute(0, 3, 1, 2)
    t = F.interpolate(t, size=(tgt, tgt), mode="bicubic", align_corners=False)
    return t.permute(0, 2, 3, 1).flatten(0, 2).to(dt)


def make_resampler_params(cfg: ResamplerConfig, seed=0, dtype=torch.float32):
    """Seeded random weights (non-trivial norms / biases so every gradient path is exercised)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in PARAM_SHAPES(cfg).items():
        if name.startswith("ln_") and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif name == "proj":
            t = cfg.embed_dim ** -0.5 * torch.randn(shape, generator=g)      # resampler.py:133-134
        elif name == "query":
            t = 0.5 * torch.randn(shape, generator=g)
        else:
            t = shape[-1] ** -0.5 * torch.randn(shape, generator=g)
        p[name] = t.to(dtype)
    p["pos_embed"] = sincos_2d(cfg.embed_dim, cfg.grid_size).to(dtype)         # frozen (resampler.py:115-118)
    return p


def resampler_forward(p, x, cfg: ResamplerConfig):
    """x [B, N, kv_dim] -> [B, num_queries, embed_dim] (resampler.py:149-168), in the dtype of the tensors."""
    B, N, _ = x.shape
    E, H, D = cfg.embed_dim, cfg.num_heads, cfg.head_dim
    pos_kv = abs_pos(p["pos_embed"], N)
    xk = x @ p["kv_proj.weight"].t()
    xn = F.layer_norm(xk, (E,), p["ln_kv.weight"], p["ln_kv.bias"], cfg.eps)
    qn = F.layer_norm(p["query"], (E,), p["ln_q.weight"], p["ln_q.bias"], cfg.eps)
    q_in = qn + p["pos_embed"]                                 # same for every image (_repeat, :165-166)
    k_in = xn + pos_kv.unsqueeze(0)
    wq, wk, wv = p["attn.in_proj_weight"].chunk(3, dim=0)
    bq, bk, bv = p["attn.in_proj_bias"].chunk(3, dim=0)
    qp = q_in @ wq.t() + bq                                    # [Q, E]
    kp = k_in @ wk.t() + bk                                    # [B, N, E]
    vp = xn @ wv.t() + bv
    qh = qp.view(-1, H, D).transpose(0, 1)                     # [H, Q, D]
    kh = kp.view(B, N, H, D).permute(0, 2, 1, 3)               # [B, H, N, D]
    vh = vp.view(B, N, H, D).permute(0, 2, 1, 3)
    s = torch.matmul(qh.unsqueeze(0) * (1.0 / math.sqrt(D)), kh.transpose(-1, -2))   # [B, H, Q, N]
    a = torch.softmax(s.float(), dim=-1).to(s.dtype)
    o = torch.matmul(a, vh).permute(0, 2, 1, 3).reshape(B, -1, E)                    # [B, Q, E]
    o = o @ p["attn.out_proj.weight"].t() + p["attn.out_proj.bias"]
    y = F.layer_norm(o, (E,), p["ln_post.weight"], p["ln_post.bias"], cfg.eps)
    return y @ p["proj"]


def synthetic_vision_tokens(cfg: ResamplerConfig, B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cfg.kv_tokens, cfg.kv_dim, generator=g)
    d_out = torch.randn(B, cfg.num_queries, cfg.embed_dim, generator=g) * 0.1
    return x, d_out
