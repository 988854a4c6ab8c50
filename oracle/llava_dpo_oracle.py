"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (plain torch tensor math, no HF modules, no kernels of this repo) of the reference's
LLaVA-1.5 DPO hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file, and only as the checker / the CPU baseline.

Parity status: PINNED against the reference itself — oracle/gen_golden.py imports the unmodified
reference code from /root/reference (llava.model.LlavaLlamaForCausalLM + transformers CLIP/Llama,
muffin collator, get_beta_and_logps, dpo_loss), runs it on seeded synthetic inputs and checks this
restatement against it (fp32, <=2e-5 rel) before writing tests/golden/*.npz.  The reference ships
no tests or golden vectors of its own.

Every function cites the reference (or pinned third-party) code it restates.  Paths are relative
to /root/reference; "HF:" = transformers/models/... (the arithmetic lives in the third-party
dependency transformers==4.35.0 pinned by pyproject.toml:22; this container has 5.5.0 whose
eager math is identical except the fp32 upcast of the logits, which is restated explicitly).

Parameters are a flat dict keyed by the HF state-dict names of LlavaLlamaForCausalLM
(model.embed_tokens.weight, model.layers.{i}.self_attn.q_proj.weight, ...,
model.mm_projector.{0,2}.{weight,bias}, model.vision_tower.vision_tower.vision_model....).
`dtype` selects fp32 ("exact" math) or bf16 (same op order as the HF bf16 eager path, so the
intermediate roundings match what the reference materialises).
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100        # llava/constants.py:8
IMAGE_TOKEN_INDEX = -200   # llava/constants.py:9


@dataclass
class OracleConfig:
    # Llama (HF: llama/configuration_llama.py)
    vocab_size: int = 32000
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 0          # 0 = num_heads (MHA); < num_heads = grouped-query attention (Mistral-7B: 32/8)
    rms_eps: float = 1e-5          # llava-v1.5-7b config.json rms_norm_eps
    rope_theta: float = 10000.0
    max_len: int = 2048            # tokenizer_model_max_length (muffin/train/train_llava15.py:249)
    # CLIP-ViT-L/14-336 (HF: clip/configuration_clip.py)
    clip_hidden: int = 1024
    clip_intermediate: int = 4096
    clip_layers: int = 24
    clip_heads: int = 16
    image_size: int = 336
    patch_size: int = 14
    clip_eps: float = 1e-5
    select_layer: int = -2         # script/train/llava15_train.sh:12 --mm_vision_select_layer -2

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads

    @property
    def kv_heads(self):
        return self.num_kv_heads or self.num_heads

    @property
    def num_patches(self):
        return (self.image_size // self.patch_size) ** 2

    @property
    def clip_layers_used(self):
        # hidden_states has clip_layers+1 entries; index select_layer (negative) => layers applied
        return self.clip_layers + 1 + self.select_layer if self.select_layer < 0 else self.select_layer


TINY = OracleConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_layers=2, num_heads=2,
                    clip_hidden=128, clip_intermediate=256, clip_layers=3, clip_heads=2,
                    image_size=56, patch_size=14)
# grouped-query decoder (4 query heads share 2 key/value heads), the attention layout of the Mistral-7B LLM inside
# OmniLMM-12B (omnilmm/model/omnilmm.py:259-265); HF Llama implements the same repeat_kv arithmetic.
TINY_GQA = OracleConfig(vocab_size=512, hidden_size=512, intermediate_size=512, num_layers=2, num_heads=4,
                        num_kv_heads=2, clip_hidden=128, clip_intermediate=256, clip_layers=3, clip_heads=2,
                        image_size=56, patch_size=14)
CONFIGS = {"TINY": TINY, "TINY_GQA": TINY_GQA}


# ------------------------------------------------------------------------------------------------
# deterministic synthetic parameters (shared by the golden generator, the oracle tests and the GPU
# parity tests so that no weight file has to be committed)
# ------------------------------------------------------------------------------------------------
def iter_params(cfg: OracleConfig, seed: int = 0, std: float = 0.02, dtype=torch.float32, scale: float = 1.0):
    """Yields (HF name, tensor) in a fixed order from ONE seeded generator, so a consumer can stream a 7B-sized
    model to the GPU tensor by tensor (tests/test_gpu_config_a.py) and still get exactly make_params' values.
    `scale` multiplies every random matrix's std (scale < 1 gives cooler logits, i.e. less bf16 noise)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(name, *shape, s=std):
        t = torch.randn(*shape, generator=g)
        t.mul_(s * scale)                      # in place: one fresh allocation per tensor (same fp32 product)
        return name, t.to(dtype)

    def ones_ish(name, n):
        return name, (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype)

    H, F_, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    yield rnd("model.embed_tokens.weight", V, H)
    for i in range(cfg.num_layers):
        pre = f"model.layers.{i}."
        Hkv = cfg.kv_heads * cfg.head_dim
        yield rnd(pre + "self_attn.q_proj.weight", H, H, s=0.05)
        yield rnd(pre + "self_attn.k_proj.weight", Hkv, H, s=0.05)
        yield rnd(pre + "self_attn.v_proj.weight", Hkv, H, s=0.05)
        yield rnd(pre + "self_attn.o_proj.weight", H, H, s=0.05)
        yield rnd(pre + "mlp.gate_proj.weight", F_, H, s=0.05)
        yield rnd(pre + "mlp.up_proj.weight", F_, H, s=0.05)
        yield rnd(pre + "mlp.down_proj.weight", H, F_, s=0.05)
        yield ones_ish(pre + "input_layernorm.weight", H)
        yield ones_ish(pre + "post_attention_layernorm.weight", H)
    yield ones_ish("model.norm.weight", H)
    yield rnd("lm_head.weight", V, H, s=0.05)
    C = cfg.clip_hidden
    yield rnd("model.mm_projector.0.weight", H, C, s=0.05)
    yield rnd("model.mm_projector.0.bias", H)
    yield rnd("model.mm_projector.2.weight", H, H, s=0.05)
    yield rnd("model.mm_projector.2.bias", H)
    vp = "model.vision_tower.vision_tower.vision_model."
    yield rnd(vp + "embeddings.class_embedding", C, s=0.5)
    yield rnd(vp + "embeddings.patch_embedding.weight", C, 3, cfg.patch_size, cfg.patch_size, s=0.05)
    yield rnd(vp + "embeddings.position_embedding.weight", cfg.num_patches + 1, C, s=0.1)
    yield ones_ish(vp + "pre_layrnorm.weight", C)
    yield rnd(vp + "pre_layrnorm.bias", C)
    for i in range(cfg.clip_layers):
        pre = vp + f"encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield rnd(pre + f"self_attn.{nm}.weight", C, C, s=0.08)
            yield rnd(pre + f"self_attn.{nm}.bias", C)
        yield rnd(pre + "mlp.fc1.weight", cfg.clip_intermediate, C, s=0.08)
        yield rnd(pre + "mlp.fc1.bias", cfg.clip_intermediate)
        yield rnd(pre + "mlp.fc2.weight", C, cfg.clip_intermediate, s=0.08)
        yield rnd(pre + "mlp.fc2.bias", C)
        yield ones_ish(pre + "layer_norm1.weight", C)
        yield rnd(pre + "layer_norm1.bias", C)
        yield ones_ish(pre + "layer_norm2.weight", C)
        yield rnd(pre + "layer_norm2.bias", C)
    yield ones_ish(vp + "post_layernorm.weight", C)   # unused by the path (select_layer=-2); kept for HF load
    yield rnd(vp + "post_layernorm.bias", C)


def make_params(cfg: OracleConfig, seed: int = 0, std: float = 0.02, dtype=torch.float32, scale: float = 1.0):
    return dict(iter_params(cfg, seed, std, dtype, scale))


def params_checksum(p):
    s = 0.0
    for k in sorted(p):
        s += float(p[k].double().abs().sum())
    return s


# ------------------------------------------------------------------------------------------------
# CLIP vision tower  (llava/model/multimodal_encoder/clip_encoder.py:36-58 -> HF CLIPVisionModel)
# ------------------------------------------------------------------------------------------------
def quick_gelu(x):
    # HF: activations.py QuickGELUActivation: x * sigmoid(1.702 x)
    return x * torch.sigmoid(1.702 * x)


def clip_features(p, images, cfg: OracleConfig):
    """images [N,3,S,S] -> hidden_states[select_layer][:, 1:]  ([N, num_patches, clip_hidden]).

    HF: clip/modeling_clip.py:202-219 (embeddings), :667-696 (pre_layrnorm + encoder),
    :354-386 (layer), :300-336 + :261-279 (attention, fp32 softmax), :339-351 (MLP, quick_gelu).
    Runs under no_grad like the reference (clip_encoder.py:46).
    """
    vp = "model.vision_tower.vision_tower.vision_model."
    dt = p[vp + "embeddings.patch_embedding.weight"].dtype
    with torch.no_grad():
        x = F.conv2d(images.to(dt), p[vp + "embeddings.patch_embedding.weight"], stride=cfg.patch_size)
        x = x.flatten(2).transpose(1, 2)                                  # [N, P, C]
        cls = p[vp + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + p[vp + "embeddings.position_embedding.weight"][None]
        x = F.layer_norm(x, (cfg.clip_hidden,), p[vp + "pre_layrnorm.weight"], p[vp + "pre_layrnorm.bias"],
                         cfg.clip_eps)
        nh = cfg.clip_heads
        hd = cfg.clip_hidden // nh
        for i in range(cfg.clip_layers_used):
            pre = vp + f"encoder.layers.{i}."
            r = x
            h = F.layer_norm(x, (cfg.clip_hidden,), p[pre + "layer_norm1.weight"], p[pre + "layer_norm1.bias"],
                             cfg.clip_eps)
            N, T, C = h.shape
            q = F.linear(h, p[pre + "self_attn.q_proj.weight"], p[pre + "self_attn.q_proj.bias"])
            k = F.linear(h, p[pre + "self_attn.k_proj.weight"], p[pre + "self_attn.k_proj.bias"])
            v = F.linear(h, p[pre + "self_attn.v_proj.weight"], p[pre + "self_attn.v_proj.bias"])
            q = q.view(N, T, nh, hd).transpose(1, 2)
            k = k.view(N, T, nh, hd).transpose(1, 2)
            v = v.view(N, T, nh, hd).transpose(1, 2)
            w = torch.matmul(q, k.transpose(-1, -2)) * (hd ** -0.5)
            w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
            a = torch.matmul(w, v).transpose(1, 2).reshape(N, T, C)
            a = F.linear(a, p[pre + "self_attn.out_proj.weight"], p[pre + "self_attn.out_proj.bias"])
            x = r + a
            r = x
            h = F.layer_norm(x, (cfg.clip_hidden,), p[pre + "layer_norm2.weight"], p[pre + "layer_norm2.bias"],
                             cfg.clip_eps)
            h = F.linear(h, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"])
            h = quick_gelu(h)
            h = F.linear(h, p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"])
            x = r + h
        return x[:, 1:].to(images.dtype if images.dtype.is_floating_point else dt)


def mm_projector(p, feats):
    """llava/model/multimodal_projector/builder.py:39-46 'mlp2x_gelu': Linear -> GELU(erf) -> Linear."""
    h = F.linear(feats, p["model.mm_projector.0.weight"], p["model.mm_projector.0.bias"])
    h = F.gelu(h)
    return F.linear(h, p["model.mm_projector.2.weight"], p["model.mm_projector.2.bias"])


# ------------------------------------------------------------------------------------------------
# image-token splice  (llava/model/llava_arch.py:150-330, attention_mask=None / labels given /
# right padding: the branch get_beta_and_logps takes, muffin/train/trainers.py:205-220)
# ------------------------------------------------------------------------------------------------
def splice_index_map(input_ids, labels, n_feat_tokens, max_len):
    """Integer part of the splice (bit-exact contract).

    Returns (src [nseq,T] int64, new_labels [nseq,T] int64, T):
      src >= 0      : position j of input_ids[b] whose token embedding fills the slot
      src = -1 - r  : row r of the flattened image-feature matrix [n_blocks*n_feat_tokens, H]
      src = INT_MIN : right padding (zeros)
    Feature blocks are consumed in batch order; a sequence without an image token consumes one
    block without emitting rows (llava_arch.py:239-246 `cur_image_idx += 1`).
    """
    PAD = -2 ** 31
    nseq, L = input_ids.shape
    rows_src, rows_lab = [], []
    blk = 0
    for b in range(nseq):
        s, l = [], []
        ids = input_ids[b].tolist()
        labs = labels[b].tolist()
        if IMAGE_TOKEN_INDEX not in ids:
            s = list(range(L))
            l = labs
            blk += 1
        else:
            for j in range(L):
                if ids[j] == IMAGE_TOKEN_INDEX:
                    s.extend(-1 - (blk * n_feat_tokens + t) for t in range(n_feat_tokens))
                    l.extend([IGNORE_INDEX] * n_feat_tokens)
                    blk += 1
                else:
                    s.append(j)
                    l.append(labs[j])
        rows_src.append(s[:max_len])
        rows_lab.append(l[:max_len])
    T = max(len(s) for s in rows_src)
    src = torch.full((nseq, T), PAD, dtype=torch.int64)
    new_labels = torch.full((nseq, T), IGNORE_INDEX, dtype=torch.int64)
    for b in range(nseq):
        n = len(rows_src[b])
        src[b, :n] = torch.tensor(rows_src[b], dtype=torch.int64)
        new_labels[b, :n] = torch.tensor(rows_lab[b], dtype=torch.int64)
    return src, new_labels, T


def splice_embeds(p, input_ids, src, image_features):
    """Row gather part of the splice: embeddings / projected image rows / zero padding (pure copies)."""
    PAD = -2 ** 31
    E = p["model.embed_tokens.weight"]
    feat = image_features.reshape(-1, image_features.shape[-1])
    nseq, T = src.shape
    out = []
    for b in range(nseq):
        s = src[b]
        tok = s >= 0
        img = (s < 0) & (s != PAD)
        ids_b = input_ids[b].clamp(min=0)
        row = torch.zeros(T, E.shape[1], dtype=E.dtype)
        row = row + 0 * feat.sum()  # keep autograd graph connected even when no image rows
        tok_idx = ids_b[s.clamp(min=0)]
        row = torch.where(tok[:, None], E[tok_idx], row)
        row = torch.where(img[:, None], feat[(-1 - s).clamp(min=0, max=feat.shape[0] - 1)], row)
        out.append(row)
    return torch.stack(out, 0)


# ------------------------------------------------------------------------------------------------
# Llama decoder  (llava/model/language_model/llava_llama.py:57-102 -> HF LlamaForCausalLM)
# ------------------------------------------------------------------------------------------------
def rms_norm(x, w, eps):
    # HF: llama/modeling_llama.py:62-67
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return w * h.to(dt)


def rope_cos_sin(T, head_dim, theta, dtype):
    # HF: llama/modeling_llama.py:124-150 (default rope, attention_scaling = 1); positions 0..T-1
    # because position_ids are not forwarded (llava_llama.py:94)
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(T, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


LORA_TARGETS = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
                "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj")   # find_all_linear_names, train_llava15_lora.py:121-134


def lora_linear(p, name, x, lora_scaling, x_adapter=None):
    """y = W x (+ (alpha/r) * B(A(x))) — peft 0.10.0 lora.Linear.forward with dropout p=0
    (muffin/train/train_llava15_lora.py:304-318: r=64, alpha=16; adapters live next to the base
    weight as `<name>.lora_A.weight` [r,in] / `<name>.lora_B.weight` [out,r]).
    PARITY: peft itself is not installed in the build container, but at dropout 0 the adapter is an exact
    reparametrisation of the base model (W' = W + s B A), so this branch is PINNED against the unmodified reference
    model run with merged weights — forward, losses and, through dL/dA = s B^T dL/dW', dL/dB = s dL/dW' A^T, every
    adapter gradient (oracle/gen_golden_lora.py, worst relative error 2e-6; fixtures tests/golden/lora/)."""
    y = F.linear(x, p[name + ".weight"])
    a = p.get(name + ".lora_A.weight")
    if a is not None:
        xa = x if x_adapter is None else x_adapter       # dropout(x) when a mask is supplied
        y = y + F.linear(F.linear(xa, a), p[name + ".lora_B.weight"]) * lora_scaling
    return y


def make_lora_params(cfg: OracleConfig, r=8, seed=3, dtype=torch.float32, b_std=0.02):
    """Adapters for every LORA_TARGETS linear. peft initialises B = 0 (adapter is a no-op at step 0);
    tests use a non-zero B so that the adapter path is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    H, F_ = cfg.hidden_size, cfg.intermediate_size
    dims = {"self_attn.q_proj": (H, H), "self_attn.k_proj": (H, H), "self_attn.v_proj": (H, H),
            "self_attn.o_proj": (H, H), "mlp.gate_proj": (F_, H), "mlp.up_proj": (F_, H), "mlp.down_proj": (H, F_)}
    out = {}
    for i in range(cfg.num_layers):
        for t in LORA_TARGETS:
            o, inn = dims[t]
            out[f"model.layers.{i}.{t}.lora_A.weight"] = (torch.randn(r, inn, generator=g) * 0.05).to(dtype)
            out[f"model.layers.{i}.{t}.lora_B.weight"] = (torch.randn(o, r, generator=g) * b_std).to(dtype)
    return out


def llama_logits(p, inputs_embeds, cfg: OracleConfig, lora_scaling=0.25, lora_drop=None):
    """inputs_embeds [nseq,T,H] -> logits [nseq,T,V] in fp32 (4.35.0 `logits.float()`).

    HF: llama/modeling_llama.py:303-333 (layer), :251-290 + :199-222 (attention: eager, causal mask
    only — attention_mask=None, muffin/train/trainers.py:199), :153-168 (RoPE), :182-184 (SwiGLU).
    """
    x = inputs_embeds
    nseq, T, H = x.shape
    nh, hd = cfg.num_heads, cfg.head_dim
    cos, sin = rope_cos_sin(T, hd, cfg.rope_theta, x.dtype)
    causal = torch.full((T, T), float("-inf")).triu(1).to(x.dtype)
    for i in range(cfg.num_layers):
        pre = f"model.layers.{i}."
        r = x
        h = rms_norm(x, p[pre + "input_layernorm.weight"], cfg.rms_eps)
        nkv = cfg.kv_heads

        def dropped(t, group):
            # lora_drop["<layer>.<group>"] = keep mask already scaled by 1/(1-p) (shape of t); None = no dropout
            m = None if lora_drop is None else lora_drop.get(f"{i}.{group}")
            return None if m is None else t * m.to(t.dtype)
        hq = dropped(h, "qkv")
        q = lora_linear(p, pre + "self_attn.q_proj", h, lora_scaling, hq).view(nseq, T, nh, hd).transpose(1, 2)
        k = lora_linear(p, pre + "self_attn.k_proj", h, lora_scaling, hq).view(nseq, T, nkv, hd).transpose(1, 2)
        v = lora_linear(p, pre + "self_attn.v_proj", h, lora_scaling, hq).view(nseq, T, nkv, hd).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        if nkv != nh:   # HF repeat_kv (llama/modeling_llama.py:171-180): kv head j serves query heads j*g .. j*g+g-1
            k = k.repeat_interleave(nh // nkv, dim=1)
            v = v.repeat_interleave(nh // nkv, dim=1)
        w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5) + causal
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        a = torch.matmul(w, v).transpose(1, 2).reshape(nseq, T, H)
        x = r + lora_linear(p, pre + "self_attn.o_proj", a, lora_scaling, dropped(a, "o"))
        r = x
        h = rms_norm(x, p[pre + "post_attention_layernorm.weight"], cfg.rms_eps)
        hg = dropped(h, "gu")
        g = lora_linear(p, pre + "mlp.gate_proj", h, lora_scaling, hg)
        u = lora_linear(p, pre + "mlp.up_proj", h, lora_scaling, hg)
        act = F.silu(g) * u
        x = r + lora_linear(p, pre + "mlp.down_proj", act, lora_scaling, dropped(act, "down"))
    x = rms_norm(x, p["model.norm.weight"], cfg.rms_eps)
    return F.linear(x, p["lm_head.weight"]).float()


# ------------------------------------------------------------------------------------------------
# log-prob gather and DPO loss
# ------------------------------------------------------------------------------------------------
def get_batch_logps(logits, labels):
    """muffin/eval/muffin_inference_logp.py:82-115 (return_all form)."""
    assert logits.shape[:-1] == labels.shape
    labels = labels[:, 1:].clone()
    logits = logits[:, :-1, :]
    loss_mask = labels != IGNORE_INDEX
    labels[labels == IGNORE_INDEX] = 0
    per_token_logps = torch.gather(logits.log_softmax(-1), dim=2, index=labels.unsqueeze(2)).squeeze(2)
    log_prob = (per_token_logps * loss_mask).sum(-1)
    average_log_prob = log_prob / loss_mask.sum(-1)
    return per_token_logps, log_prob, average_log_prob


def dpo_loss(policy_chosen_logps, policy_rejected_logps, reference_chosen_logps, reference_rejected_logps, beta):
    """muffin/train/trainers.py:91-126."""
    pi_logratios = policy_chosen_logps - policy_rejected_logps
    ref_logratios = reference_chosen_logps - reference_rejected_logps
    logits = pi_logratios - ref_logratios
    losses = -F.logsigmoid(beta * logits)
    chosen_rewards = beta * (policy_chosen_logps - reference_chosen_logps).detach()
    rejected_rewards = beta * (policy_rejected_logps - reference_rejected_logps).detach()
    return losses, chosen_rewards, rejected_rewards


def policy_logps(p, cfg, concatenated_input_ids, concatenated_labels, images, dedup_images=True, lora_drop=None):
    """get_beta_and_logps' llava15 branch (muffin/train/trainers.py:185-231): images are repeated for
    the win and rej halves (identical copies, so encoding them once is exact)."""
    feats = clip_features(p, images, cfg)                       # [B, P, C]  (no grad)
    proj = mm_projector(p, feats.to(p["model.mm_projector.0.weight"].dtype))
    image_features = torch.cat([proj, proj], dim=0)             # trainers.py:190
    src, new_labels, T = splice_index_map(concatenated_input_ids, concatenated_labels, cfg.num_patches,
                                          cfg.max_len)
    embeds = splice_embeds(p, concatenated_input_ids, src, image_features)
    logits = llama_logits(p, embeds, cfg, lora_drop=lora_drop)
    per_tok, logp, avg = get_batch_logps(logits, new_labels)
    return dict(src=src, labels=new_labels, per_token_logps=per_tok, logp=logp, avg_logp=avg,
                inputs_embeds=embeds, logits=logits)


def dpo_step(p, cfg, batch, beta=0.1, dpo_weight=1.0, sft_weight=0.0, use_average=False):
    """LLaVA15DPOTrainer.compute_loss (muffin/train/trainers.py:279-311) minus logging.
    batch: concatenated_input_ids/labels [2B,L] (win first), images [B,3,S,S], ref_win_logp, ref_rej_logp [B]."""
    out = policy_logps(p, cfg, batch["concatenated_input_ids"], batch["concatenated_labels"], batch["images"])
    lp = out["avg_logp"] if use_average else out["logp"]
    B = lp.shape[0] // 2
    pw, pr = lp[:B], lp[B:]
    losses, cr, rr = dpo_loss(pw, pr, batch["ref_win_logp"], batch["ref_rej_logp"], beta)
    loss = dpo_weight * losses.mean() - sft_weight * pw.mean()
    out.update(policy_win_logp=pw, policy_rej_logp=pr, losses=losses, chosen_rewards=cr, rejected_rewards=rr,
               reward_accuracies=(cr > rr).float(), loss=loss)
    return out


TRAINABLE_PREFIXES = ("model.embed_tokens.", "model.layers.", "model.norm.", "lm_head.", "model.mm_projector.")


def trainable_names(p):
    return [k for k in p if k.startswith(TRAINABLE_PREFIXES)]


def adamw_update(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01):
    """torch.optim.AdamW single-tensor math (optim=adamw_torch, muffin/train/train_llava15.py:75;
    wd/lr from script/train/llava15_train.sh:31-32), fp32 master weights as under DeepSpeed bf16."""
    param = param * (1 - lr * wd)
    exp_avg = beta1 * exp_avg + (1 - beta1) * grad
    exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) * grad * grad
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = exp_avg_sq.sqrt() / math.sqrt(bc2) + eps
    param = param - (lr / bc1) * exp_avg / denom
    return param, exp_avg, exp_avg_sq


# ------------------------------------------------------------------------------------------------
# synthetic batch  (SURVEY.md §8d canonical inputs)
# ------------------------------------------------------------------------------------------------
def synthetic_pair_batch(cfg, B, prompt_len, resp_len, seed, image_pos=None, ragged=False):
    g = torch.Generator().manual_seed(seed)
    image_pos = prompt_len - 13 if image_pos is None else image_pos
    ids_w, ids_r, lab_w, lab_r = [], [], [], []
    for _ in range(B):
        prompt = torch.randint(3, cfg.vocab_size, (prompt_len,), generator=g)
        prompt[0] = 1
        prompt[image_pos] = IMAGE_TOKEN_INDEX
        for dst_i, dst_l in ((ids_w, lab_w), (ids_r, lab_r)):
            R = resp_len if not ragged else int(torch.randint(resp_len // 2, resp_len + 1, (1,), generator=g))
            resp = torch.randint(3, cfg.vocab_size, (R,), generator=g)
            resp[-1] = 2
            dst_i.append(torch.cat([prompt, resp]))
            dst_l.append(torch.cat([torch.full((prompt_len,), IGNORE_INDEX), resp]))
    L = max(len(x) for x in ids_w + ids_r)

    def pad(rows, value):
        return torch.stack([torch.cat([r, torch.full((L - len(r),), value, dtype=torch.int64)]) for r in rows])

    # preference_collator_fn order: win first, then rej; pad id 0, label pad -100
    # (muffin/eval/muffin_inference_logp.py:187-208, muffin/train/train_utils.py:55-96)
    input_ids = torch.cat([pad(ids_w, 0), pad(ids_r, 0)], 0)
    labels = torch.cat([pad(lab_w, IGNORE_INDEX), pad(lab_r, IGNORE_INDEX)], 0)
    images = torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=g)
    return dict(concatenated_input_ids=input_ids, concatenated_labels=labels, images=images)
