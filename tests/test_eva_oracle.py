"""CPU: the restated EVA vision tower (oracle/eva_oracle.py; timm 0.9.10's eva02_enormous_patch14_clip_224 as the
reference instantiates it, omnilmm/model/omnilmm.py:31-52). timm is absent, so nothing here is a pin against the
library — these tests hold the restatement to the properties the reference's call sites fix, and prove the two
identities the CUDA tower relies on (head padding, position-resample matrix)."""
import torch

from oracle import eva_oracle as E


def test_shapes_follow_the_reference_call_sites():
    cfg = E.EvaConfig()
    # omnilmm.py:47-51: the resampler's kv_dim is vision_tower.embed_dim; 448 px / 14 = 32 x 32 patch tokens
    assert cfg.embed_dim == 1792 and cfg.head_dim == 112 and cfg.live_blocks == 63 and cfg.grid ** 2 == 1024
    t = E.TINY_EVA
    p = E.make_eva_params(t, seed=1)
    img = torch.randn(2, 3, t.img_size, t.img_size, generator=torch.Generator().manual_seed(0))
    feats = E.eva_forward_features(p, img, t)
    assert feats.shape == (2, 1 + t.grid ** 2, t.embed_dim)
    assert E.vision_tokens(p, img, t).shape == (2, t.grid ** 2, t.embed_dim)       # prefix token dropped (omnilmm.py:115-117)
    assert not any(k.startswith("blocks.%d." % t.live_blocks) for k in p)           # blocks[-1] = Identity (omnilmm.py:43)


def test_pos_resample_matrix_is_the_interpolation():
    t = E.TINY_EVA
    R = E.pos_resample_matrix(t)
    pos = torch.randn(1, 1 + t.pretrain_grid ** 2, 8, generator=torch.Generator().manual_seed(3))
    want = E.resample_abs_pos_embed(pos, (t.grid, t.grid), 1)[0]
    assert torch.allclose(R @ pos[0], want, atol=1e-5)
    assert torch.equal(R[0], torch.eye(R.shape[1])[0])                               # prefix row passes through
    assert torch.allclose(R[1:, 1:].sum(-1), torch.ones(t.grid ** 2), atol=1e-5)     # interpolation weights sum to 1
    from rlaifv_b200.eva_tower import EvaDims, pos_resample_matrix
    d = EvaDims(embed_dim=t.embed_dim, depth=t.depth, num_heads=t.num_heads, mlp_hidden=t.mlp_hidden,
                pretrain_img=t.pretrain_img, img_size=t.img_size)
    assert torch.allclose(pos_resample_matrix(d), R, atol=1e-6)                      # host constant of the CUDA module


def test_zero_padding_heads_to_128_changes_nothing():
    """The CUDA tower stores each 112-wide head padded to 128 (zero qkv rows / proj columns): same output, and the
    gradients of the padded elements are exactly zero."""
    t = E.TINY_EVA
    p = E.make_eva_params(t, seed=2)
    g = torch.Generator().manual_seed(4)
    B, N, C, nh, hd, HP = 2, 10, t.embed_dim, t.num_heads, t.head_dim, 128
    x = torch.randn(B, N, C, generator=g)
    W, Wo = p["blocks.0.attn.qkv.weight"], p["blocks.0.attn.proj.weight"]
    qkv = torch.nn.functional.linear(x, W).reshape(B, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    ref = torch.nn.functional.linear((((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(-1) @ v).transpose(1, 2)
                                     .reshape(B, N, C), Wo)
    Wp = torch.zeros(3, nh, HP, C)
    Wp[:, :, :hd] = W.view(3, nh, hd, C)
    Wp = Wp.reshape(3 * nh * HP, C).requires_grad_(True)
    Wop = torch.zeros(C, nh, HP)
    Wop[:, :, :hd] = Wo.view(C, nh, hd)
    Wop = Wop.reshape(C, nh * HP).requires_grad_(True)
    qkv = torch.nn.functional.linear(x, Wp).reshape(B, N, 3, nh, HP).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    got = torch.nn.functional.linear((((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(-1) @ v).transpose(1, 2)
                                     .reshape(B, N, nh * HP), Wop)
    assert torch.allclose(got, ref, atol=1e-5)
    got.square().sum().backward()
    assert float(Wp.grad.view(3, nh, HP, C)[:, :, hd:].abs().max()) == 0.0
    assert float(Wop.grad.view(C, nh, HP)[:, :, hd:].abs().max()) == 0.0


def test_k_bias_is_not_a_parameter():
    # timm's EvaAttention registers k_bias as a zero BUFFER: only q_bias / v_bias are trainable
    p = E.make_eva_params(E.TINY_EVA)
    assert "blocks.0.attn.q_bias" in p and "blocks.0.attn.v_bias" in p and "blocks.0.attn.k_bias" not in p
