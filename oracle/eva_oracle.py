"""TEST INFRASTRUCTURE ONLY — CPU restatement of OmniLMM-12B's vision tower (BASELINE config d).
Only tests/ (and oracle tooling) may import this.

PARITY UNPINNED.  The arithmetic lives in a third-party dependency that is absent from /root/reference and from this
image: timm==0.9.10 (pyproject.toml:21).  The reference only names the model
(`timm.create_model('eva02_enormous_patch14_clip_224.laion2b_plus', pretrained=False, num_classes=0,
dynamic_img_size=True, dynamic_img_pad=True)`, omnilmm/model/omnilmm.py:31-36), replaces its last block by Identity
(:43) and consumes `forward_features(pixel_values)[:, num_prefix_tokens:]` (omnilmm.py:107-120).  There is no timm
source, wheel or fixture to pin against here, so this file restates timm 0.9.10's published algorithm
(timm/models/eva.py) and anchors on the reference's call sites:

  eva02_enormous_patch14_clip_224      eva.py model_args: img_size=224, patch_size=14, embed_dim=1792, depth=64,
                                       num_heads=16 (head_dim 112), mlp_ratio=15360/1792, use_post_norm=True,
                                       global_pool='token'.  ("A EVA-CLIP specific variant that uses residual post-norm
                                       in blocks": plain GELU MLP, absolute position embedding, no RoPE, no SwiGLU,
                                       no sub-LN — those belong to the other eva02_* sizes.)
  Eva.forward_features                 patch_embed -> _pos_embed -> blocks -> self.norm (LayerNorm, because
                                       global_pool='token' => use_fc_norm False)
  PatchEmbed (dynamic_img_size)        Conv2d(3, C, 14, stride 14, bias) -> NHWC tokens
  Eva._pos_embed (dynamic_img_size)    pos_embed [1, 1+16*16, C] resampled to the input grid by
                                       resample_abs_pos_embed (bicubic, antialias, align_corners=False, computed in
                                       fp32; the prefix token's row is kept), cls_token prepended, x + pos_embed
  EvaBlockPostNorm.forward             x = x + norm1(attn(x));  x = x + norm2(mlp(x))
  EvaAttention (qkv_fused, qkv_bias)   qkv = F.linear(x, qkv.weight, cat(q_bias, zeros (k_bias buffer), v_bias));
                                       softmax(q k^T * head_dim^-0.5) v; proj (bias)
  Mlp                                  fc1 (bias) -> nn.GELU (erf) -> fc2 (bias)
  timm.layers.LayerNorm                eps = 1e-6

`blocks[-1] = Identity()` leaves depth-1 = 63 live blocks; 448 px inputs give 32x32 = 1024 patch tokens + cls.
"""
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class EvaConfig:
    embed_dim: int = 1792
    depth: int = 64                 # as created; the LAST block is replaced by Identity (omnilmm.py:43)
    num_heads: int = 16
    mlp_hidden: int = 15360
    patch_size: int = 14
    pretrain_img: int = 224         # pos_embed is defined on (pretrain_img / patch)^2 positions
    img_size: int = 448             # what the OmniLMM pipeline feeds (build_transform input_size)
    eps: float = 1e-6

    @property
    def head_dim(self):
        return self.embed_dim // self.num_heads

    @property
    def live_blocks(self):
        return self.depth - 1

    @property
    def grid(self):
        return self.img_size // self.patch_size

    @property
    def pretrain_grid(self):
        return self.pretrain_img // self.patch_size


# tiny tower for parity tests: 2 heads of width 112 (the real head_dim — the CUDA path pads it to 128), 3 live
# blocks, 4x4 pre-training grid resampled to 6x6
TINY_EVA = EvaConfig(embed_dim=224, depth=4, num_heads=2, mlp_hidden=448, patch_size=14, pretrain_img=56, img_size=84)


def make_eva_params(cfg: EvaConfig, seed=0, scale=1.0, dtype=torch.float32):
    """timm state-dict names of the Eva model (live blocks only)."""
    g = torch.Generator().manual_seed(seed)
    C, Hd = cfg.embed_dim, cfg.mlp_hidden
    p = {}

    def rnd(name, *shape, s=0.02):
        p[name] = (torch.randn(*shape, generator=g) * (s * scale)).to(dtype)

    def ones_ish(name, n):
        p[name] = (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype)

    rnd("patch_embed.proj.weight", C, 3, cfg.patch_size, cfg.patch_size, s=0.05)
    rnd("patch_embed.proj.bias", C)
    rnd("cls_token", 1, 1, C, s=0.3)
    rnd("pos_embed", 1, 1 + cfg.pretrain_grid ** 2, C, s=0.1)
    for i in range(cfg.live_blocks):
        pre = f"blocks.{i}."
        rnd(pre + "attn.qkv.weight", 3 * C, C, s=0.08)
        rnd(pre + "attn.q_bias", C)
        rnd(pre + "attn.v_bias", C)
        rnd(pre + "attn.proj.weight", C, C, s=0.08)
        rnd(pre + "attn.proj.bias", C)
        ones_ish(pre + "norm1.weight", C)
        rnd(pre + "norm1.bias", C)
        rnd(pre + "mlp.fc1.weight", Hd, C, s=0.08)
        rnd(pre + "mlp.fc1.bias", Hd)
        rnd(pre + "mlp.fc2.weight", C, Hd, s=0.08)
        rnd(pre + "mlp.fc2.bias", C)
        ones_ish(pre + "norm2.weight", C)
        rnd(pre + "norm2.bias", C)
    ones_ish("norm.weight", C)
    rnd("norm.bias", C)
    return p


def resample_abs_pos_embed(posemb, new_hw, num_prefix_tokens=1):
    """timm.layers.pos_embed.resample_abs_pos_embed (bicubic, antialias=True, fp32 compute, cast back)."""
    n_new = new_hw[0] * new_hw[1] + num_prefix_tokens
    if n_new == posemb.shape[1] and new_hw[0] == new_hw[1]:
        return posemb
    prefix, grid = posemb[:, :num_prefix_tokens], posemb[:, num_prefix_tokens:]
    old = int(round(grid.shape[1] ** 0.5))
    dt = grid.dtype
    g = grid.float().reshape(1, old, old, -1).permute(0, 3, 1, 2)
    g = F.interpolate(g, size=new_hw, mode="bicubic", antialias=True, align_corners=False)
    g = g.permute(0, 2, 3, 1).reshape(1, -1, grid.shape[-1]).to(dt)
    return torch.cat([prefix, g], dim=1)


def pos_resample_matrix(cfg: EvaConfig):
    """The same resampling as a dense matrix R [1 + grid^2, 1 + pretrain_grid^2] (prefix row passes through): the
    interpolation is linear in pos_embed, so pos_resampled = R @ pos_embed. Host-side constant of the CUDA module."""
    n_old = 1 + cfg.pretrain_grid ** 2
    eye = torch.eye(n_old).unsqueeze(0)                       # [1, n_old, n_old]: "channel" c = basis vector c
    out = resample_abs_pos_embed(eye, (cfg.grid, cfg.grid), 1)   # [1, n_new, n_old]
    return out[0]


def eva_forward_features(p, images, cfg: EvaConfig):
    """images [B, 3, S, S] -> [B, 1 + grid^2, C] (Eva.forward_features with the last block = Identity)."""
    dt = p["pos_embed"].dtype
    C, nh, hd = cfg.embed_dim, cfg.num_heads, cfg.head_dim
    x = F.conv2d(images.to(dt), p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=cfg.patch_size)
    B, _, H, W = x.shape
    x = x.permute(0, 2, 3, 1).reshape(B, H * W, C)
    pos = resample_abs_pos_embed(p["pos_embed"], (H, W), 1)
    x = torch.cat([p["cls_token"].expand(B, -1, -1), x], dim=1) + pos
    N = x.shape[1]
    zeros = torch.zeros(C, dtype=dt)
    for i in range(cfg.live_blocks):
        pre = f"blocks.{i}."
        qkv_bias = torch.cat([p[pre + "attn.q_bias"], zeros, p[pre + "attn.v_bias"]])
        qkv = F.linear(x, p[pre + "attn.qkv.weight"], qkv_bias).reshape(B, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        att = (q * hd ** -0.5) @ k.transpose(-2, -1)
        att = att.softmax(dim=-1)
        a = (att @ v).transpose(1, 2).reshape(B, N, C)
        a = F.linear(a, p[pre + "attn.proj.weight"], p[pre + "attn.proj.bias"])
        x = x + F.layer_norm(a, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], cfg.eps)
        m = F.linear(x, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"])
        m = F.linear(F.gelu(m), p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"])
        x = x + F.layer_norm(m, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], cfg.eps)
    return F.layer_norm(x, (C,), p["norm.weight"], p["norm.bias"], cfg.eps)


def vision_tokens(p, images, cfg: EvaConfig):
    """OmniLMMModel.get_vision_embedding up to the resampler (omnilmm.py:113-118): prefix (cls) token dropped."""
    return eva_forward_features(p, images, cfg)[:, 1:]
