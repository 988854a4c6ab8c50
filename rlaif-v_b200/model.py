"""LLaVA-1.5 DPO policy on the B200 kernels: parameter storage + hand-written forward / backward.

Mirrors what the reference executes in `get_beta_and_logps(..., is_llava15=True)`
(muffin/train/trainers.py:205-231): CLIP tower (frozen, llava/model/multimodal_encoder/clip_encoder.py:46-58)
-> mm_projector (llava/model/multimodal_projector/builder.py:39-46) -> image-token splice
(llava/model/llava_arch.py:150-330) -> Llama decoder (llava/model/language_model/llava_llama.py:57-102)
-> per-token log-prob gather (muffin/eval/muffin_inference_logp.py:82-115) — and the backward of
all trainable parts.  There is no autograd here: the backward is an explicit sequence of kernel
launches over an activation stash.

All arithmetic runs in the C-ABI library (ops.*); torch supplies device memory, streams and
zero-fills.  Parameters live in ONE flat bf16 buffer cut into buckets (embed | layer i | head |
projector); gradients in a same-layout flat buffer, so the ZeRO-2 optimizer (zero2.py) can
reduce-scatter / update / all-gather bucket slices in place.
"""
import math
from dataclasses import dataclass, field

import torch

from . import ops

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
_BF = torch.bfloat16
_F32 = torch.float32


@dataclass
class LlavaDims:
    """Dimensions (defaults = LLaVA-1.5-7B + CLIP-ViT-L/14-336; script/train/llava15_train.sh:8,11)."""
    vocab_size: int = 32000
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 0          # 0 = num_heads; fewer = grouped-query attention (Mistral-7B: 32 / 8)
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_len: int = 2048
    clip_hidden: int = 1024
    clip_intermediate: int = 4096
    clip_layers: int = 24
    clip_heads: int = 16
    image_size: int = 336
    patch_size: int = 14
    clip_eps: float = 1e-5
    select_layer: int = -2
    # vision front-end: "clip_mlp" = frozen CLIP tower + trainable mlp2x_gelu projector (LLaVA-1.5);
    # "resampler" = OmniLMM-12B: external vision-tower tokens -> trainable perceiver resampler -> in-place
    # <im_patch> splice (omnilmm_model.OmniLMMDPOPolicy; omnilmm/model/omnilmm.py:107-120, 183-265)
    frontend: str = "clip_mlp"
    num_query: int = 64            # resampler queries = <im_patch> tokens per image (omnilmm.py:46-51)
    vision_width: int = 1792       # EVA-02-E token width
    im_patch_token: int = -1
    im_start_token: int = -1
    im_end_token: int = -1

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads

    @property
    def kv_heads(self):
        return self.num_kv_heads or self.num_heads

    @property
    def kv_size(self):
        return self.kv_heads * self.head_dim

    @property
    def num_patches(self):
        return (self.image_size // self.patch_size) ** 2

    @property
    def clip_layers_used(self):
        return self.clip_layers + 1 + self.select_layer if self.select_layer < 0 else self.select_layer

    @property
    def patch_k(self):
        return 3 * self.patch_size * self.patch_size

    @property
    def patch_k_pad(self):
        return (self.patch_k + 63) // 64 * 64


def _round_up(n, m):
    return (n + m - 1) // m * m


@dataclass
class Segment:
    name: str           # storage name (fused tensors have their own names)
    shape: tuple
    offset: int         # element offset in the flat buffer
    decay: bool


@dataclass
class Bucket:
    name: str
    start: int
    size: int            # padded to a multiple of PAD
    decay_size: int      # leading part that gets weight decay; the rest (norm weights, biases) does not
    segments: list = field(default_factory=list)


class ParamStore:
    """Flat bf16 parameter / gradient storage with named views.

    Storage tensors: embed, per layer {qkv [3H,H], o [H,H], gu [2F,H], down [H,F], ln1 [H], ln2 [H]},
    norm [H], lm_head [V,H], projector {w0 [H,C], w2 [H,H], b0 [H], b2 [H]}.
    HF state-dict names map onto (views of) these — see hf_views().
    """
    PAD = 1024  # bucket sizes are multiples of this (divisible by any world size <= 8 with 128-elem slices)

    def __init__(self, dims: LlavaDims, device):
        self.dims = dims
        d = dims
        H, F, V, C = d.hidden_size, d.intermediate_size, d.vocab_size, d.clip_hidden
        KV = d.kv_size
        self.buckets = []
        off = 0

        def add_bucket(name, decay_list, nodecay_list):
            nonlocal off
            b = Bucket(name=name, start=off, size=0, decay_size=0)
            o = off
            for nm, shape in decay_list:
                b.segments.append(Segment(nm, shape, o, True))
                o += math.prod(shape)
            b.decay_size = o - off
            for nm, shape in nodecay_list:
                b.segments.append(Segment(nm, shape, o, False))
                o += math.prod(shape)
            b.size = _round_up(o - off, self.PAD)
            off += b.size
            self.buckets.append(b)

        add_bucket("embed", [("embed", (V, H))], [])
        for i in range(d.num_layers):
            add_bucket(f"layer{i}",
                       [(f"l{i}.qkv", (H + 2 * KV, H)), (f"l{i}.o", (H, H)), (f"l{i}.gu", (2 * F, H)),
                        (f"l{i}.down", (H, F))],
                       [(f"l{i}.ln1", (H,)), (f"l{i}.ln2", (H,))])
        add_bucket("head", [("lm_head", (V, H))], [("norm", (H,))])
        if d.frontend == "clip_mlp":
            add_bucket("projector", [("proj.w0", (H, C)), ("proj.w2", (H, H))], [("proj.b0", (H,)), ("proj.b2", (H,))])
        self.numel = off
        self.flat = torch.zeros(off, dtype=_BF, device=device)
        self.grad = torch.zeros(off, dtype=_BF, device=device)
        self.p = {}
        self.g = {}
        for b in self.buckets:
            for s in b.segments:
                n = math.prod(s.shape)
                self.p[s.name] = self.flat[s.offset:s.offset + n].view(*s.shape)
                self.g[s.name] = self.grad[s.offset:s.offset + n].view(*s.shape)

    def hf_views(self):
        """HF state-dict name -> view into the fused storage (LlavaLlamaForCausalLM naming)."""
        d = self.dims
        H, F, KV = d.hidden_size, d.intermediate_size, d.kv_size
        out = {"model.embed_tokens.weight": self.p["embed"], "model.norm.weight": self.p["norm"],
               "lm_head.weight": self.p["lm_head"]}
        if d.frontend == "clip_mlp":
            out.update({"model.mm_projector.0.weight": self.p["proj.w0"], "model.mm_projector.0.bias": self.p["proj.b0"],
                        "model.mm_projector.2.weight": self.p["proj.w2"], "model.mm_projector.2.bias": self.p["proj.b2"]})
        for i in range(d.num_layers):
            pre = f"model.layers.{i}."
            qkv, gu = self.p[f"l{i}.qkv"], self.p[f"l{i}.gu"]
            out[pre + "self_attn.q_proj.weight"] = qkv[0:H]
            out[pre + "self_attn.k_proj.weight"] = qkv[H:H + KV]
            out[pre + "self_attn.v_proj.weight"] = qkv[H + KV:H + 2 * KV]
            out[pre + "self_attn.o_proj.weight"] = self.p[f"l{i}.o"]
            out[pre + "mlp.gate_proj.weight"] = gu[0:F]
            out[pre + "mlp.up_proj.weight"] = gu[F:2 * F]
            out[pre + "mlp.down_proj.weight"] = self.p[f"l{i}.down"]
            out[pre + "input_layernorm.weight"] = self.p[f"l{i}.ln1"]
            out[pre + "post_attention_layernorm.weight"] = self.p[f"l{i}.ln2"]
        return out

    def hf_grad_views(self):
        saved = self.p
        self.p = self.g
        try:
            return self.hf_views()
        finally:
            self.p = saved

    def load_hf(self, state):
        """Copy a (possibly fp32, CPU) HF-named state dict into the flat storage."""
        views = self.hf_views()
        for k, v in views.items():
            v.copy_(state[k].to(device=v.device, dtype=_BF))



class LoraStore:
    """Flat bf16 storage of the LoRA adapters (config e; muffin/train/train_llava15_lora.py:112-134, 304-318:
    r=64, alpha=16 on q,k,v,o,gate,up,down of every decoder layer; lm_head / projector / vision excluded).
    One bucket per layer: A_qkv [3r,H] (the three lora_A stacked — q,k,v share their input), B_q/B_k/B_v [H,r],
    A_o [r,H], B_o [H,r], A_gu [2r,H], B_g/B_u [F,r], A_d [r,F], B_d [H,r]. Adapters get weight decay like any
    other matrix (HF decay groups exclude only norms and biases)."""

    def __init__(self, dims: LlavaDims, device, r=64, alpha=16, seed=7, init_b_zero=True, dropout=0.0):
        self.dims, self.r, self.scaling = dims, r, alpha / r
        self.dropout = float(dropout)      # applied to the adapter input while training (peft lora_dropout)
        self.rng_seed = 0x5EED0000 + seed
        self.step = 0                      # advanced by the engine once per optimisation step
        H, F = dims.hidden_size, dims.intermediate_size
        self.buckets = []
        off = 0
        shapes = [("A_qkv", (3 * r, H)), ("B_q", (H, r)), ("B_k", (H, r)), ("B_v", (H, r)), ("A_o", (r, H)),
                  ("B_o", (H, r)), ("A_gu", (2 * r, H)), ("B_g", (F, r)), ("B_u", (F, r)), ("A_d", (r, F)),
                  ("B_d", (H, r))]
        for i in range(dims.num_layers):
            b = Bucket(name=f"lora{i}", start=off, size=0, decay_size=0)
            o = off
            for nm, shape in shapes:
                b.segments.append(Segment(f"l{i}.{nm}", shape, o, True))
                o += math.prod(shape)
            b.decay_size = o - off
            b.size = _round_up(o - off, ParamStore.PAD)
            off += b.size
            self.buckets.append(b)
        self.numel = off
        self.flat = torch.zeros(off, dtype=_BF, device=device)
        self.grad = torch.zeros(off, dtype=_BF, device=device)
        self.p, self.g = {}, {}
        for b in self.buckets:
            for sg in b.segments:
                n = math.prod(sg.shape)
                self.p[sg.name] = self.flat[sg.offset:sg.offset + n].view(*sg.shape)
                self.g[sg.name] = self.grad[sg.offset:sg.offset + n].view(*sg.shape)
        for i in range(dims.num_layers):   # stacked lora_B views (adjacent segments) = second-source B operand
            for src in (self.p, self.g):
                q, gq = src[f"l{i}.B_q"], src[f"l{i}.B_g"]
                base = self.flat if src is self.p else self.grad
                oq = (q.data_ptr() - base.data_ptr()) // 2
                og = (gq.data_ptr() - base.data_ptr()) // 2
                src[f"l{i}.B_qkv"] = base[oq:oq + 3 * H * r].view(3 * H, r)
                src[f"l{i}.B_gu"] = base[og:og + 2 * F * r].view(2 * F, r)
                src[f"l{i}.B_o_cat"] = src[f"l{i}.B_o"]
                src[f"l{i}.B_d_cat"] = src[f"l{i}.B_d"]
        # peft init: lora_A kaiming_uniform(a=sqrt(5)) = U(-1/sqrt(in), 1/sqrt(in)), lora_B = 0
        g = torch.Generator(device=device).manual_seed(seed)
        for name, v in list(self.p.items()):
            if name.endswith(("B_qkv", "B_gu", "B_o_cat", "B_d_cat")):
                continue
            if ".A_" in name:
                bound = 1.0 / math.sqrt(v.shape[1])
                v.copy_((torch.rand(v.shape, generator=g, device=device) * 2 - 1) * bound)
            elif not init_b_zero:
                v.copy_(torch.randn(v.shape, generator=g, device=device) * 0.02)

    _PEFT = {"q_proj": ("A_qkv", 0, "B_q"), "k_proj": ("A_qkv", 1, "B_k"), "v_proj": ("A_qkv", 2, "B_v"),
             "o_proj": ("A_o", 0, "B_o"), "gate_proj": ("A_gu", 0, "B_g"), "up_proj": ("A_gu", 1, "B_u"),
             "down_proj": ("A_d", 0, "B_d")}

    def hf_views(self, grads=False):
        """`model.layers.{i}.{self_attn|mlp}.{x}_proj.lora_{A,B}.weight` -> views (peft adapter naming minus the
        `base_model.model.` prefix and the adapter name)."""
        src = self.g if grads else self.p
        r, out = self.r, {}
        for i in range(self.dims.num_layers):
            for proj, (a_name, idx, b_name) in self._PEFT.items():
                mod = "self_attn" if proj in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
                pre = f"model.layers.{i}.{mod}.{proj}."
                out[pre + "lora_A.weight"] = src[f"l{i}.{a_name}"][idx * r:(idx + 1) * r]
                out[pre + "lora_B.weight"] = src[f"l{i}.{b_name}"]
        return out

    def load_hf(self, state):
        for k, v in self.hf_views().items():
            v.copy_(state[k].to(device=v.device, dtype=_BF))


class ClipWeights:
    """Frozen CLIP-ViT weights in kernel-friendly form (fused qkv, padded patch-embedding matrix)."""

    def __init__(self, dims: LlavaDims, device, state=None, seed=1):
        d = dims
        C, I = d.clip_hidden, d.clip_intermediate
        vp = "model.vision_tower.vision_tower.vision_model."
        if state is None:
            g = torch.Generator().manual_seed(seed)
            state = {}

            def rnd(name, *shape, s=0.02):
                state[name] = torch.randn(*shape, generator=g) * s

            rnd(vp + "embeddings.class_embedding", C)
            rnd(vp + "embeddings.patch_embedding.weight", C, 3, d.patch_size, d.patch_size)
            rnd(vp + "embeddings.position_embedding.weight", d.num_patches + 1, C)
            state[vp + "pre_layrnorm.weight"] = torch.ones(C)
            state[vp + "pre_layrnorm.bias"] = torch.zeros(C)
            for i in range(d.clip_layers_used):
                pre = vp + f"encoder.layers.{i}."
                for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
                    rnd(pre + f"self_attn.{nm}.weight", C, C)
                    state[pre + f"self_attn.{nm}.bias"] = torch.zeros(C)
                rnd(pre + "mlp.fc1.weight", I, C)
                state[pre + "mlp.fc1.bias"] = torch.zeros(I)
                rnd(pre + "mlp.fc2.weight", C, I)
                state[pre + "mlp.fc2.bias"] = torch.zeros(C)
                for ln in ("layer_norm1", "layer_norm2"):
                    state[pre + ln + ".weight"] = torch.ones(C)
                    state[pre + ln + ".bias"] = torch.zeros(C)

        def dev(t):
            return t.to(device=device, dtype=_BF).contiguous()

        self.cls = dev(state[vp + "embeddings.class_embedding"])
        w = state[vp + "embeddings.patch_embedding.weight"].reshape(C, -1).to(_BF)
        wp = torch.zeros(C, d.patch_k_pad, dtype=_BF)
        wp[:, : d.patch_k] = w
        self.patch_w = wp.to(device)
        self.pos = dev(state[vp + "embeddings.position_embedding.weight"])
        self.pre_ln_w = dev(state[vp + "pre_layrnorm.weight"])
        self.pre_ln_b = dev(state[vp + "pre_layrnorm.bias"])
        self.layers = []
        for i in range(d.clip_layers_used):
            pre = vp + f"encoder.layers.{i}."
            L = {}
            L["qkv_w"] = dev(torch.cat([state[pre + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")], 0))
            L["qkv_b"] = dev(torch.cat([state[pre + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")], 0))
            L["o_w"] = dev(state[pre + "self_attn.out_proj.weight"])
            L["o_b"] = dev(state[pre + "self_attn.out_proj.bias"])
            L["fc1_w"] = dev(state[pre + "mlp.fc1.weight"])
            L["fc1_b"] = dev(state[pre + "mlp.fc1.bias"])
            L["fc2_w"] = dev(state[pre + "mlp.fc2.weight"])
            L["fc2_b"] = dev(state[pre + "mlp.fc2.bias"])
            for ln in ("layer_norm1", "layer_norm2"):
                L[ln + "_w"] = dev(state[pre + ln + ".weight"])
                L[ln + "_b"] = dev(state[pre + ln + ".bias"])
            self.layers.append(L)


class LlavaDPOPolicy:
    """Forward (policy log-probs) and backward (gradients into ParamStore.grad) of one micro-batch."""

    def __init__(self, dims: LlavaDims, device="cuda", hf_state=None, seed=0, init_std=0.02):
        self.dims = dims
        self.device = torch.device(device)
        self.store = ParamStore(dims, self.device)
        has_clip = dims.frontend == "clip_mlp"
        if hf_state is not None:
            self.store.load_hf(hf_state)
            self.clip = ClipWeights(dims, self.device, hf_state) if has_clip else None
        else:
            self._random_init(seed, init_std)
            self.clip = ClipWeights(dims, self.device, None, seed + 1) if has_clip else None
        self._rope = {}
        self._bufs = {}
        # True: also stash the normalised inputs and the SwiGLU product (22 KB/token/layer more memory,
        # 3 fewer row passes per layer in the backward). The engine turns it on when HBM allows.
        self.stash_extra = False
        # middle ground for one GPU (unsharded 81 GB optimizer state): stash only the SwiGLU product (22 KB/token/layer,
        # 12.8 GB at config b) — saves the largest of the three recompute passes
        self.stash_act = False
        # training forward sends only the supervised positions through final norm / lm_head / log-softmax
        # (forward_logps); inference / the reference log-prob pre-pass keep the full head (per-token values of every
        # position are part of the parquet contract)
        self.compact_head = True
        # attention backward as two kernels (dK/dV, dQ) without fp32 atomics; False = the fused round-1 kernel
        self.split_attention_bwd = True
        # SwiGLU backward inside the epilogue of the down-projection dgrad (no d(act) round trip, one launch less per
        # layer); False = separate swiglu_bwd pass
        self.fuse_swiglu_bwd = True
        self.embed_grad_f32 = None   # fp32 scatter target for embedding rows (allocated lazily)
        self._stash = None
        self.lora = None             # LoraStore: base weights frozen, adapters + mm_projector trainable

    def enable_lora(self, r=64, alpha=16, seed=7, init_b_zero=True, dropout=0.0):
        if self.dims.kv_heads != self.dims.num_heads:
            raise NotImplementedError("LoRA adapters on a grouped-query decoder (unequal q/k/v widths)")
        self.lora = LoraStore(self.dims, self.device, r=r, alpha=alpha, seed=seed, init_b_zero=init_b_zero,
                              dropout=dropout)
        return self.lora

    training = True

    def _lora_seed(self, i, group):
        """One dropout stream per (optimisation step, micro-batch forward, layer, linear group)."""
        L = self.lora
        g = ("qkv", "o", "gu", "down").index(group)
        return (L.rng_seed * 1000003 + L.step) * 4099 + (self._fwd_count * 131 + i) * 4 + g

    _fwd_count = 0

    # ---- what the optimizer trains / in which order the forward needs it ----
    def trainable_buckets(self):
        from .zero2 import store_buckets
        if self.lora is None:
            return store_buckets(self.store)
        return store_buckets(self.lora) + store_buckets(self.store, names={"projector"})

    def param_need_order(self):
        d = self.dims
        if self.lora is None:
            return ["projector", "embed"] + [f"layer{i}" for i in range(d.num_layers)] + ["head"]
        return ["projector"] + [f"lora{i}" for i in range(d.num_layers)]

    def layer_bucket_name(self, i):
        return f"layer{i}" if self.lora is None else f"lora{i}"

    def tail_bucket_names(self):
        """Buckets whose gradients are final only after the whole backward (reduced at the end of the step)."""
        return ["embed", "projector"]

    # ---- one linear group = base GEMM (+ LoRA adapters sharing the input) ----
    _GROUPS = {"qkv": ("A_qkv", ("B_q", "B_k", "B_v"), "B_qkv"), "o": ("A_o", ("B_o",), "B_o_cat"),
               "gu": ("A_gu", ("B_g", "B_u"), "B_gu"), "down": ("A_d", ("B_d",), "B_d_cat")}

    def _lin_fwd(self, i, group, x, out, residual=None, ls=None):
        """out = x @ W^T (+ residual). LoRA: t = s * x @ A^T (all sub-linears' lora_A stacked), then ONE GEMM
        accumulates x @ W^T and t_j @ B_j^T in the same TMEM tile (rlaifv_gemm_bf16_dual)."""
        W = self.store.p[f"l{i}.{group}"]
        if self.lora is None:
            return ops.gemm(x, W, out, residual=residual)
        L = self.lora
        a_name, b_names, bcat = self._GROUPS[group]
        A = L.p[f"l{i}.{a_name}"]
        M = x.shape[0]
        t = torch.empty((M, A.shape[0]), dtype=_BF, device=self.device) if ls is not None else \
            self.buf("lora_t_" + group, (M, A.shape[0]))
        xa = x
        if L.dropout > 0.0 and self.training and ls is not None:
            # the sub-linears of a fused group share one mask (peft draws one per adapter; same marginal law)
            seed = self._lora_seed(i, group)
            xa = ops.dropout_fwd(x, L.dropout, seed, out=self.buf("lora_xd_%d" % x.shape[1], tuple(x.shape)))
            ls["seed_" + group] = seed
        ops.gemm(xa, A, t, alpha=L.scaling)
        n_sub = out.shape[1] // len(b_names) if len(b_names) > 1 else 0
        ops.gemm_dual(x, W, t, L.p[f"l{i}.{bcat}"], out, k2=L.r, r=L.r, n_sub=n_sub, residual=residual)
        if ls is not None:
            ls["t_" + group] = t
        return out

    def _lin_bwd(self, i, group, dy, x, dx_out, ls, acc, glu=None):
        """dx = dy @ W (+ adapter path); weight gradients: base (full FT) or adapters (LoRA).
        glu (down projection only): the layer's gate|up pre-activations [M, 2F]; dx_out is then [M, 2F] and receives
        swiglu_bwd(glu, dy @ W) from the GEMM epilogue — d(act) stays in the accumulator (ops.gemm_swiglu_bwd)."""
        W = self.store.p[f"l{i}.{group}"]
        if self.lora is None:
            ops.gemm(dy, x, self.store.g[f"l{i}.{group}"], a_mn=True, b_mn=True, accumulate=acc)
            if glu is not None:
                return ops.gemm_swiglu_bwd(dy, W, glu, dx_out)
            return ops.gemm(dy, W, dx_out, b_mn=True)
        L = self.lora
        a_name, b_names, _ = self._GROUPS[group]
        A = L.p[f"l{i}.{a_name}"]
        r, M = L.r, x.shape[0]
        t = ls["t_" + group]                      # already scaled by s
        dt = self.buf("lora_dt_" + group, (M, A.shape[0]))
        n_out = dy.shape[1] // len(b_names)
        for j, bn in enumerate(b_names):
            dy_j = dy[:, j * n_out:(j + 1) * n_out]
            ops.gemm(dy_j, L.p[f"l{i}.{bn}"], dt[:, j * r:(j + 1) * r], b_mn=True, alpha=L.scaling)   # dt_j = s dy_j B_j
            ops.gemm(dy_j, t[:, j * r:(j + 1) * r], L.g[f"l{i}.{bn}"], a_mn=True, b_mn=True, accumulate=acc)  # dB_j
        seed = ls.get("seed_" + group)
        if seed is not None:
            # dropout on the adapter input: dA sees the dropped input, and the adapter's dx passes through the mask
            xd = ops.dropout_fwd(x, L.dropout, seed, out=self.buf("lora_xd_%d" % x.shape[1], tuple(x.shape)))
            ops.gemm(dt, xd, L.g[f"l{i}.{a_name}"], a_mn=True, b_mn=True, accumulate=acc)             # dA = dt^T drop(x)
            dx_plain = dx_out if glu is None else self.buf("dact", tuple(x.shape))
            ops.gemm(dy, W, dx_plain, b_mn=True)
            pa = ops.gemm(dt, A, self.buf("lora_xd_%d" % x.shape[1], tuple(x.shape)), b_mn=True)                       # dt @ A
            ops.dropout_bwd_add(dx_plain, pa, L.dropout, seed)
            return dx_plain if glu is None else ops.swiglu_bwd(glu, dx_plain, dx_out)
        ops.gemm(dt, x, L.g[f"l{i}.{a_name}"], a_mn=True, b_mn=True, accumulate=acc)                  # dA = dt^T x
        # dx = dy @ W + dt @ A in one pass (second source = stacked lora_A, MN-major like W)
        if glu is not None:
            return ops.gemm_swiglu_bwd(dy, W, glu, dx_out, a2=dt, b2=A)
        return ops.gemm_dual(dy, W, dt, A, dx_out, k2=A.shape[0], r=r, n_sub=0, b_mn=True)

    # ------------------------------------------------------------------ init / buffers
    def _random_init(self, seed, std):
        g = torch.Generator(device=self.device).manual_seed(seed)
        for b in self.store.buckets:
            for s in b.segments:
                v = self.store.p[s.name]
                if len(s.shape) == 1 and (s.name.endswith("ln1") or s.name.endswith("ln2") or s.name == "norm"):
                    v.fill_(1.0)
                elif len(s.shape) == 1:
                    v.zero_()
                else:
                    # chunked normal_ to bound the fp32 temporary
                    flat = v.view(-1)
                    step = 1 << 26
                    for o in range(0, flat.numel(), step):
                        n = min(step, flat.numel() - o)
                        flat[o:o + n].copy_(torch.randn(n, generator=g, device=self.device, dtype=_F32) * std)

    def buf(self, key, shape, dtype=_BF):
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def rope_tables(self, T):
        if T not in self._rope:
            d = self.dims
            hd = d.head_dim
            inv_freq = 1.0 / (d.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
            pos = torch.arange(T, dtype=_F32)
            emb = torch.cat([pos[:, None] * inv_freq[None, :]] * 2, dim=-1)
            self._rope[T] = (emb.cos().to(_BF).to(self.device).contiguous(),
                             emb.sin().to(_BF).to(self.device).contiguous())
        return self._rope[T]

    # ------------------------------------------------------------------ CLIP (frozen, forward only)
    def encode_images(self, images):
        """images [b,3,S,S] (any float dtype) -> CLIP patch features [b*P, C] (bf16)."""
        d, cw = self.dims, self.clip
        b = images.shape[0]
        P, C = d.num_patches, d.clip_hidden
        T = P + 1
        img = images.to(device=self.device, dtype=_BF).contiguous()
        cols = ops.clip_im2col(img, d.patch_size, d.patch_k_pad)
        patch = ops.gemm(cols, cw.patch_w)
        x = ops.clip_embed(patch, cw.cls, cw.pos, b, P)
        x = ops.layernorm_fwd(x, cw.pre_ln_w, cw.pre_ln_b, d.clip_eps)
        nh = d.clip_heads
        hd = C // nh
        scale = hd ** -0.5
        for L in cw.layers:
            h = ops.layernorm_fwd(x, L["layer_norm1_w"], L["layer_norm1_b"], d.clip_eps, out=self.buf("clip_h", x.shape))
            qkv = ops.gemm(h, L["qkv_w"], self.buf("clip_qkv", (b * T, 3 * C)), bias=L["qkv_b"])
            att, _ = ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], b, T, nh, hd, False, scale,
                                       out=self.buf("clip_att", (b * T, C)), lse=self.buf("clip_lse", (b, nh, T), _F32))
            x2 = ops.gemm(att, L["o_w"], self.buf("clip_x2", x.shape), bias=L["o_b"], residual=x)
            h = ops.layernorm_fwd(x2, L["layer_norm2_w"], L["layer_norm2_b"], d.clip_eps, out=self.buf("clip_h", x.shape))
            f = ops.gemm(h, L["fc1_w"], self.buf("clip_f", (b * T, d.clip_intermediate)), bias=L["fc1_b"],
                         act=ops.ACT_QUICK_GELU)
            x = ops.gemm(f, L["fc2_w"], self.buf("clip_x", x.shape), bias=L["fc2_b"], residual=x2)
        return ops.clip_drop_cls(x, b, P)

    # ------------------------------------------------------------------ splice
    def splice(self, input_ids, labels, image_rows, n_blocks, img_index=None, T_hint=None):
        """llava_arch.prepare_inputs_labels_for_multimodal (attention_mask=None path).

        image_rows: [n_blocks*P, H] projected features; img_index[i] = feature block used by the
        i-th image slot (default: identity).  Returns (embeds [nseq*T,H], new_labels [nseq,T], src, T)."""
        d = self.dims
        nseq, L = input_ids.shape
        n_img, lens = ops.splice_count(input_ids, d.num_patches, d.max_len)
        if T_hint is None:
            T = int(lens.max().item())       # one tiny D2H per step (the reference syncs ~4x per sample)
            n_slots = int(torch.clamp(n_img, min=1).sum().item())
        else:
            T, n_slots = T_hint
        if img_index is None:
            img_index = torch.arange(n_slots, dtype=torch.int32, device=self.device)
        src, new_labels = ops.splice_map(input_ids, labels, n_img, img_index, d.num_patches, T, d.max_len)
        embeds = ops.splice_gather(src, input_ids, self.store.p["embed"], image_rows,
                                   out=self.buf("x0", (nseq * T, d.hidden_size)))
        return embeds, new_labels, src, T

    # ------------------------------------------------------------------ decoder layers
    def _run_layers(self, x, nseq, T, st):
        """32 x [RMSNorm -> fused qkv GEMM -> RoPE -> causal attention -> o GEMM(+res) -> RMSNorm ->
        fused gate|up GEMM -> SwiGLU -> down GEMM(+res)]; st (dict or None) receives the per-layer stash."""
        d, P, dev = self.dims, self.store.p, self.device
        H, F, KV = d.hidden_size, d.intermediate_size, d.kv_size
        nh, hd, nkv = d.num_heads, d.head_dim, d.kv_heads
        QKV = H + 2 * KV
        M = nseq * T
        cos, sin = self.rope_tables(T)
        scale = hd ** -0.5
        keep_stash = st is not None
        for i in range(d.num_layers):
            self._need(self.layer_bucket_name(i))
            ls = None
            if keep_stash:
                ls = {"x": x}
                qkv = torch.empty((M, QKV), dtype=_BF, device=dev)
                att = torch.empty((M, H), dtype=_BF, device=dev)
                x2 = torch.empty((M, H), dtype=_BF, device=dev)
                gu = torch.empty((M, 2 * F), dtype=_BF, device=dev)
                x3 = torch.empty((M, H), dtype=_BF, device=dev)
                rstd1 = torch.empty(M, dtype=_F32, device=dev)
                rstd2 = torch.empty(M, dtype=_F32, device=dev)
                lse = torch.empty((nseq, nh, T), dtype=_F32, device=dev)
            else:
                qkv, att = self.buf("qkv", (M, QKV)), self.buf("att", (M, H))
                x2, gu = self.buf("x2", (M, H)), self.buf("gu", (M, 2 * F))
                x3 = self.buf("x3_%d" % (i & 1), (M, H))
                rstd1 = rstd2 = None
                lse = self.buf("lse", (nseq, nh, T), _F32)
            extra = keep_stash and self.stash_extra
            n1 = ops.rmsnorm_fwd(x, P[f"l{i}.ln1"], d.rms_eps,
                                 out=torch.empty((M, H), dtype=_BF, device=dev) if extra else self.buf("n", (M, H)),
                                 rstd=rstd1)
            self._lin_fwd(i, "qkv", n1, qkv, ls=ls if keep_stash else None)
            ops.rope_fwd(qkv, cos, sin, T, nh, hd, n_kv_heads=nkv)
            ops.attention_fwd(qkv[:, :H], qkv[:, H:H + KV], qkv[:, H + KV:], nseq, T, nh, hd, True, scale, out=att,
                              lse=lse, n_kv_heads=nkv)
            self._lin_fwd(i, "o", att, x2, residual=x, ls=ls if keep_stash else None)
            n2 = ops.rmsnorm_fwd(x2, P[f"l{i}.ln2"], d.rms_eps,
                                 out=torch.empty((M, H), dtype=_BF, device=dev) if extra else self.buf("n", (M, H)),
                                 rstd=rstd2)
            self._lin_fwd(i, "gu", n2, gu, ls=ls if keep_stash else None)
            keep_act = extra or (keep_stash and self.stash_act)
            act = ops.swiglu_fwd(gu, torch.empty((M, F), dtype=_BF, device=dev) if keep_act else self.buf("act", (M, F)))
            self._lin_fwd(i, "down", act, x3, residual=x2, ls=ls if keep_stash else None)
            if keep_stash:
                ls.update(qkv=qkv, att=att, x2=x2, gu=gu, rstd1=rstd1, rstd2=rstd2, lse=lse)
                if extra:
                    ls.update(n1=n1, n2=n2, act=act)
                elif keep_act:
                    ls.update(act=act)
                st["layers"].append(ls)
            x = x3
        return x

    def decoder_logits(self, inputs_embeds):
        """inputs_embeds [nseq,T,H] bf16 -> logits [nseq,T,V] bf16 (inference form, no stash)."""
        d = self.dims
        nseq, T, H = inputs_embeds.shape
        x = inputs_embeds.to(device=self.device, dtype=_BF).reshape(nseq * T, H).contiguous()
        x = self._run_layers(x, nseq, T, None)
        self._need("head")
        hn = ops.rmsnorm_fwd(x, self.store.p["norm"], d.rms_eps, out=self.buf("hn", (nseq * T, H)))
        logits = ops.gemm(hn, self.store.p["lm_head"])
        return logits.view(nseq, T, d.vocab_size)

    # ------------------------------------------------------------------ forward
    def forward_logps(self, input_ids, labels, images, keep_stash=True, return_per_token=True, T_hint=None):
        """input_ids/labels [2b, L] int64 (win rows first, then rej — preference_collator_fn order),
        images [b,3,S,S].  Returns dict(per_token_logps [2b,T-1], logp [2b], avg_logp [2b], labels [2b,T])."""
        d, P = self.dims, self.store.p
        dev = self.device
        input_ids = input_ids.to(dev).contiguous()
        labels = labels.to(dev).contiguous()
        nseq = input_ids.shape[0]
        b = images.shape[0]
        H, F, V = d.hidden_size, d.intermediate_size, d.vocab_size
        nh, hd = d.num_heads, d.head_dim
        st = {"layers": []} if keep_stash else None
        self._fwd_count += 1

        proj = self._frontend_fwd(images, st)                                # [b*Pn, H] image rows
        # images are shared by the win and rej copy of a pair (trainers.py:190 cat([images, images]))
        n_slots = nseq
        img_index = (torch.arange(n_slots, dtype=torch.int32, device=dev) % b).contiguous() if nseq == 2 * b \
            else torch.arange(n_slots, dtype=torch.int32, device=dev)
        self._need("embed")
        x, new_labels, src, T = self.splice(input_ids, labels, proj, b, img_index,
                                            T_hint=(T_hint, n_slots) if T_hint is not None else None)
        M = nseq * T
        if keep_stash:
            st.update(src=src, input_ids=input_ids, labels=new_labels, T=T, nseq=nseq, b=b,
                      n_feat_rows=proj.shape[0])
        x = self._run_layers(x, nseq, T, st)
        self._need("head")
        compact = keep_stash and self.compact_head
        if compact:
            # training: only the positions whose next token is supervised reach the final norm / lm_head / log-softmax
            # (get_batch_logps masks the rest out of the sum, muffin_inference_logp.py:93-104). A row of the collator
            # holds at most L labels, so L slots per sequence always suffice — no host sync for the count.
            row_pos = ops.supervised_rows(new_labels, input_ids.shape[1])
            R = row_pos.numel()
            xc = ops.rows_gather(row_pos, x, out=self.buf("xc", (R, H)))
            rstd_f = torch.empty(R, dtype=_F32, device=dev)
            hn = ops.rmsnorm_fwd(xc, P["norm"], d.rms_eps, out=self.buf("hn_c", (R, H)), rstd=rstd_f)
            logits = ops.gemm(hn, P["lm_head"], self.buf("logits_c", (R, V)))
            per_tok, lse_v, logp, avg, count = ops.logp_fwd_rows(logits, new_labels, row_pos, nseq, T)
            st.update(x_final=xc, rstd_f=rstd_f, hn=hn, logits=logits, lse_v=lse_v, count=count, row_pos=row_pos)
            self._stash = st
            return dict(per_token_logps=per_tok, logp=logp, avg_logp=avg, labels=new_labels, T=T)
        rstd_f = torch.empty(M, dtype=_F32, device=dev) if keep_stash else None
        hn = ops.rmsnorm_fwd(x, P["norm"], d.rms_eps, out=self.buf("hn", (M, H)), rstd=rstd_f)
        logits = ops.gemm(hn, P["lm_head"], self.buf("logits", (M, V)))
        per_tok, lse_v, logp, avg, count = ops.logp_fwd(logits, new_labels, nseq, T)
        if keep_stash:
            st.update(x_final=x, rstd_f=rstd_f, hn=hn, logits=logits, lse_v=lse_v, count=count)
            self._stash = st
        return dict(per_token_logps=per_tok, logp=logp, avg_logp=avg, labels=new_labels, T=T)

    # ------------------------------------------------------------------ backward
    def backward_logps(self, d_logp, use_average=False, accumulate=False, token_weight=None, weight_sum=None):
        """Gradients of sum_b d_logp[b] * logp[b] into store.grad (bf16; `accumulate` adds to what is
        there — used for the 2nd.. micro-batch of a step). token_weight [nseq, T-1] fp32: logp is the token-weighted
        sum of compute_weighted_logp (muffin/train/trainers.py:128-137); with use_average also pass weight_sum."""
        st = self._stash
        assert st is not None, "forward_logps(keep_stash=True) must precede backward_logps"
        d, P, G = self.dims, self.store.p, self.store.g
        dev = self.device
        nseq, T, b = st["nseq"], st["T"], st["b"]
        M = nseq * T
        H, F, V, KV = d.hidden_size, d.intermediate_size, d.vocab_size, d.kv_size
        nh, hd, nkv = d.num_heads, d.head_dim, d.kv_heads
        scale = hd ** -0.5
        cos, sin = self.rope_tables(T)
        acc = bool(accumulate)

        row_pos = st.get("row_pos")
        if token_weight is not None:
            assert not use_average or weight_sum is not None
        if row_pos is not None:                   # compact head (training): logits hold the supervised rows only
            norm = (weight_sum if token_weight is not None else st["count"]) if use_average else None
            dlogits = ops.logp_bwd_rows(st["logits"], st["labels"], row_pos, st["lse_v"], d_logp, nseq, T,
                                        token_weight=token_weight, norm=norm)
        elif token_weight is not None:
            dlogits = ops.logp_bwd_weighted(st["logits"], st["labels"], st["lse_v"], d_logp, token_weight, nseq, T,
                                            wsum=weight_sum if use_average else None)
        else:
            dlogits = ops.logp_bwd(st["logits"], st["labels"], st["lse_v"], d_logp, nseq, T,
                                   count=st["count"] if use_average else None)
        frozen = self.lora is not None            # LoRA: lm_head / norms / embeddings / base matrices are frozen
        scratch_h = self.buf("frozen_dw", (H,)) if frozen else None
        Mh = dlogits.shape[0]                     # rows that went through the head (M, or the compact row count)
        if not frozen:
            ops.gemm(dlogits, st["hn"], G["lm_head"], a_mn=True, b_mn=True, accumulate=acc)    # dW = dlogits^T hn
        dhn = ops.gemm(dlogits, P["lm_head"], self.buf("dn_c" if row_pos is not None else "dn", (Mh, H)), b_mn=True)               # dhn = dlogits W
        dx = ops.rmsnorm_bwd(dhn, st["x_final"], P["norm"], st["rstd_f"],
                             self.buf("dx_c" if row_pos is not None else "dx_a", (Mh, H)),
                             scratch_h if frozen else G["norm"], dw_accumulate=acc and not frozen)
        if row_pos is not None:                   # back to the [M, H] residual stream (zeros at unsupervised positions)
            full = self.buf("dx_a", (M, H))
            full.zero_()
            dx = ops.rows_scatter(row_pos, dx, full)
        if self.on_head_grads_ready is not None:
            self.on_head_grads_ready()
        for i in reversed(range(d.num_layers)):
            ls = st["layers"][i]
            # ---- MLP ----
            act = ls["act"] if "act" in ls else ops.swiglu_fwd(ls["gu"], self.buf("act", (M, F)))   # recompute
            if self.fuse_swiglu_bwd and F % 64 == 0:
                # d(act) = dx @ W_down never leaves the GEMM: its epilogue applies the SwiGLU backward
                dgu = self._lin_bwd(i, "down", dx, act, self.buf("dgu", (M, 2 * F)), ls, acc, glu=ls["gu"])
            else:
                dact = self._lin_bwd(i, "down", dx, act, self.buf("dact", (M, F)), ls, acc)
                dgu = ops.swiglu_bwd(ls["gu"], dact, self.buf("dgu", (M, 2 * F)))
            n2 = ls["n2"] if "n2" in ls else ops.rmsnorm_fwd(ls["x2"], P[f"l{i}.ln2"], d.rms_eps,
                                                             out=self.buf("n", (M, H)))             # recompute
            dn2 = self._lin_bwd(i, "gu", dgu, n2, self.buf("dn", (M, H)), ls, acc)
            dx2 = ops.rmsnorm_bwd(dn2, ls["x2"], P[f"l{i}.ln2"], ls["rstd2"], self.buf("dx_b", (M, H)),
                                  scratch_h if frozen else G[f"l{i}.ln2"], dres=dx, dw_accumulate=acc and not frozen)
            # ---- attention ----
            datt = self._lin_bwd(i, "o", dx2, ls["att"], self.buf("datt", (M, H)), ls, acc)
            qkv = ls["qkv"]
            dqkv = self.buf("dqkv", (M, H + 2 * KV))
            if self.split_attention_bwd:
                # dK/dV kernel + dQ kernel: dQ lands in the q block of dqkv as bf16, rotated back in place
                ops.attention_bwd_split(qkv[:, :H], qkv[:, H:H + KV], qkv[:, H + KV:], ls["att"], datt, ls["lse"], nseq,
                                        T, T, nh, hd, True, scale, dqkv[:, :H], dqkv[:, H:H + KV], dqkv[:, H + KV:],
                                        n_kv_heads=nkv, delta_ws=self.buf("delta", (nseq, nh, T), _F32))
                ops.rope_bwd(dqkv, None, cos, sin, T, nh, hd, n_kv_heads=nkv)
            else:
                dq32 = self.buf("dq32", (M, H), _F32)
                dq32.zero_()
                ops.attention_bwd(qkv[:, :H], qkv[:, H:H + KV], qkv[:, H + KV:], ls["att"], datt, ls["lse"], nseq, T, nh,
                                  hd, scale, dq32, dqkv[:, H:H + KV], dqkv[:, H + KV:],
                                  self.buf("delta", (nseq, nh, T), _F32), n_kv_heads=nkv)
                ops.rope_bwd(dqkv, dq32, cos, sin, T, nh, hd, n_kv_heads=nkv)
            n1 = ls["n1"] if "n1" in ls else ops.rmsnorm_fwd(ls["x"], P[f"l{i}.ln1"], d.rms_eps,
                                                             out=self.buf("n", (M, H)))             # recompute
            dn1 = self._lin_bwd(i, "qkv", dqkv, n1, self.buf("dn", (M, H)), ls, acc)
            dx = ops.rmsnorm_bwd(dn1, ls["x"], P[f"l{i}.ln1"], ls["rstd1"], self.buf("dx_a", (M, H)),
                                 scratch_h if frozen else G[f"l{i}.ln1"], dres=dx2, dw_accumulate=acc and not frozen)
            st["layers"][i] = None   # release this layer's stash
            if self.on_layer_grads_ready is not None:
                self.on_layer_grads_ready(i)
        # ---- splice backward: embedding rows + projected image rows ----
        if frozen:
            pass
        elif self.embed_grad_f32 is None:
            self.embed_grad_f32 = torch.zeros((V, H), dtype=_F32, device=dev)
        elif not acc:
            self.embed_grad_f32.zero_()
        n_feat_rows = st["n_feat_rows"]
        dfeat32 = self.buf("dfeat32", (n_feat_rows, H), _F32)
        dfeat32.zero_()
        ops.splice_scatter(st["src"], st["input_ids"], dx, None if frozen else self.embed_grad_f32, dfeat32)
        dproj = ops.f32_to_bf16(dfeat32, self.buf("dproj", (n_feat_rows, H)))
        self._frontend_bwd(dproj, st, acc)
        self._stash = None

    # ------------------------------------------------------------------ vision front-end (LLaVA-1.5: CLIP + projector)
    def _frontend_fwd(self, images, st):
        """images [b,3,S,S] -> projected image rows [b*P, H]; st (dict or None) receives what the backward needs."""
        P, H = self.store.p, self.dims.hidden_size
        feats = self.encode_images(images)                                   # [b*Pn, C]
        self._need("projector")
        pre = ops.gemm(feats, P["proj.w0"], self.buf("proj_pre", (feats.shape[0], H)), bias=P["proj.b0"])
        post = ops.gelu_fwd(pre, self.buf("proj_post", pre.shape))
        proj = ops.gemm(post, P["proj.w2"], self.buf("proj_out", pre.shape), bias=P["proj.b2"])
        if st is not None:
            st.update(feats=feats, proj_pre=pre, proj_post=post)
        return proj

    def _frontend_bwd(self, dproj, st, acc):
        """dproj [b*P, H] bf16 = gradient of the image rows -> mm_projector gradients (the CLIP tower is frozen)."""
        P, G, H = self.store.p, self.store.g, self.dims.hidden_size
        n_feat_rows = dproj.shape[0]
        ops.gemm(dproj, st["proj_post"], G["proj.w2"], a_mn=True, b_mn=True, accumulate=acc)
        ops.colsum(dproj, G["proj.b2"], accumulate=acc)
        dpost = ops.gemm(dproj, P["proj.w2"], self.buf("dpost", (n_feat_rows, H)), b_mn=True)
        dpre = ops.gelu_bwd(st["proj_pre"], dpost, self.buf("dpre", (n_feat_rows, H)))
        ops.gemm(dpre, st["feats"], G["proj.w0"], a_mn=True, b_mn=True, accumulate=acc)
        ops.colsum(dpre, G["proj.b0"], accumulate=acc)

    on_layer_grads_ready = None
    on_head_grads_ready = None
    param_ready = None          # callable(bucket_name): wait for that bucket's parameter all-gather (ZeRO-2)

    def _need(self, bucket):
        """Order the current stream after `bucket`'s ZeRO-2 parameter all-gather. Frozen buckets (LoRA: embed / head /
        base layers) are not the optimizer's and need no wait; an unknown trainable name is a bug and raises."""
        if self.param_ready is None:
            return
        if self.lora is not None and not (bucket.startswith("lora") or bucket in ("projector", "resampler")):
            return
        self.param_ready(bucket)

    def finalize_embed_grad(self):
        """fp32 embedding-row accumulator -> bf16 flat gradient (once per optimizer step)."""
        if self.lora is not None:
            return
        ops.f32_to_bf16(self.embed_grad_f32.view(-1), self.store.g["embed"].view(-1))
