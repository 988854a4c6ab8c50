"""-m gpu: OmniLMM's vision tower on the CUDA library (rlaif-v_b200/eva_tower.py) — forward and
backward against the restated timm model (oracle/eva_oracle.py, parity-unpinned: timm is absent), then the WHOLE
config-(d) policy (tower -> resampler -> in-place splice -> GQA decoder -> log-probs -> DPO loss -> backward -> AdamW)
against the composed oracle (tower restatement + the pinned OmniLMM oracle)."""
import pytest
import torch

from oracle import eva_oracle as E
from oracle import omnilmm_oracle as OM
from oracle import resampler_oracle as R

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def tower_dims(t=E.TINY_EVA):
    from rlaifv_b200.eva_tower import EvaDims
    return EvaDims(embed_dim=t.embed_dim, depth=t.depth, num_heads=t.num_heads, mlp_hidden=t.mlp_hidden,
                   patch_size=t.patch_size, pretrain_img=t.pretrain_img, img_size=t.img_size, eps=t.eps)


def test_tower_forward_backward_match_restated_timm_model():
    from rlaifv_b200.eva_tower import EvaTower
    t = E.TINY_EVA
    params = E.make_eva_params(t, seed=5)
    tower = EvaTower(tower_dims(t), "cuda", state=params)
    rt = tower.timm_state()
    for k, v in params.items():                                  # padded storage round-trips the timm state exactly
        assert torch.equal(rt[k].float().cpu(), v.to(torch.bfloat16).float()), k
    g = torch.Generator().manual_seed(6)
    B = 3
    img = torch.randn(B, 3, t.img_size, t.img_size, generator=g)
    tok = tower.forward(img, keep_stash=True)
    pf = {k: v.to(torch.bfloat16).float().requires_grad_(True) for k, v in params.items()}   # same bf16-rounded weights
    ref = E.vision_tokens(pf, img.to(torch.bfloat16).float(), t)
    e_fwd = float((tok.float().cpu() - ref.detach()).abs().mean() / ref.detach().abs().mean())
    print(f"tower tokens: mean-abs rel err {e_fwd:.2e}, max rel {rel(tok.float(), ref.detach()):.2e}")
    assert e_fwd <= 5e-3 and rel(tok.float(), ref.detach()) <= 4e-2
    d_tok = torch.randn(ref.shape, generator=g) * 0.1
    (ref * d_tok).sum().backward()
    tower.backward(d_tok.cuda())
    torch.cuda.synchronize()
    got = tower.timm_state(grads=True)
    worst = 0.0
    for k in params:
        gr, gg = pf[k].grad.double(), got[k].double().cpu().view_as(pf[k].grad)
        err = float((gg - gr).norm() / (gr.norm() + 1e-30))
        worst = max(worst, err)
        assert err <= 4e-2, (k, err)
    print(f"tower parameter gradients ({len(params)} tensors): worst relative L2 error {worst:.2e}")
    # padded head elements never receive gradient; the (untrained) k-bias slots stay zero
    d = tower.dims
    qw = tower.g["b0.qkv_w"].view(3, d.num_heads, 128, d.embed_dim)
    assert float(qw[:, :, d.head_dim:].float().abs().max()) == 0.0
    assert float(tower.g["b0.qkv_b"].view(3, d.num_heads, 128)[1].float().abs().max()) == 0.0
    assert float(tower.g["b0.proj_w"].view(d.embed_dim, d.num_heads, 128)[:, :, d.head_dim:].float().abs().max()) == 0.0


def full_setup(B=2, seed=3):
    from rlaifv_b200.model import LlavaDims
    from rlaifv_b200.omnilmm_model import OmniLMMDPOPolicy
    t, dec, tok = E.TINY_EVA, OM.TINY_OMNI_DEC, OM.TINY_OMNI_TOK
    res = R.ResamplerConfig(grid_size=4, embed_dim=dec.hidden_size, num_heads=dec.hidden_size // 128,
                            kv_dim=t.embed_dim, kv_tokens=t.grid ** 2)
    params = OM.make_omnilmm_params(dec, res, seed)
    params.update({"model.vision_tower." + k: v for k, v in E.make_eva_params(t, seed=seed + 9).items()})
    dims = LlavaDims(frontend="resampler", vocab_size=dec.vocab_size, hidden_size=dec.hidden_size,
                     intermediate_size=dec.intermediate_size, num_layers=dec.num_layers, num_heads=dec.num_heads,
                     num_kv_heads=dec.num_kv_heads, rms_eps=dec.rms_eps, num_query=res.num_queries,
                     vision_width=res.kv_dim, im_patch_token=tok.im_patch, im_start_token=tok.im_start,
                     im_end_token=tok.im_end)
    pol = OmniLMMDPOPolicy(dims, "cuda", hf_state=params, eva_dims=tower_dims(t))
    batch = OM.synthetic_omni_batch(dec, res, tok, B, 28, 20, seed=seed + 7)
    g = torch.Generator().manual_seed(seed + 11)
    batch["images"] = torch.randn(B, 3, t.img_size, t.img_size, generator=g)
    return pol, params, batch, (t, dec, res, tok)


def test_whole_omnilmm_policy_with_tower_matches_composed_oracle():
    from rlaifv_b200 import ops
    pol, params, batch, (t, dec, res, tok) = full_setup()
    B = batch["images"].shape[0]
    ids, labels = batch["concatenated_input_ids"], batch["concatenated_labels"]
    out = pol.forward_logps(ids, labels, batch["images"], keep_stash=True)
    pf = {k: v.to(torch.bfloat16).float().requires_grad_(True) for k, v in params.items()}
    tp = {k[len("model.vision_tower."):]: v for k, v in pf.items() if k.startswith("model.vision_tower.")}
    vt = E.vision_tokens(tp, batch["images"].to(torch.bfloat16).float(), t)
    rw, rr = torch.tensor([-60.0, -62.0]), torch.tensor([-61.0, -60.5])
    ob = dict(concatenated_input_ids=ids, concatenated_labels=labels, vision_tokens=vt, ref_win_logp=rw, ref_rej_logp=rr)
    oo = OM.omnilmm_dpo_step(pf, dec, res, tok, ob, beta=0.1)
    e_sum = rel(out["logp"], oo["logp"].detach())
    print(f"summed log-probs (pixels -> tower -> ... -> gather) rel err vs the composed fp32 oracle {e_sum:.2e}")
    assert e_sum <= 2e-3
    oo["loss"].backward()
    _, _, _, dpw, dpr, out9 = ops.dpo_loss(out["logp"][:B].contiguous(), out["logp"][B:].contiguous(), rw.cuda(),
                                           rr.cuda(), 0.1)
    pol.backward_logps(torch.cat([dpw, dpr]).contiguous())
    pol.finalize_embed_grad()
    torch.cuda.synchronize()
    assert abs(float(out9[0]) - float(oo["loss"])) <= 5e-3 * max(1.0, abs(float(oo["loss"])))
    got = pol.tower.timm_state(grads=True)
    worst = 0.0
    for k, v in tp.items():
        gr = v.grad.double()
        err = float((got[k].double().cpu().view_as(gr) - gr).norm() / (gr.norm() + 1e-30))
        worst = max(worst, err)
        assert err <= 6e-2, (k, err)
    print(f"tower gradients through resampler + decoder: worst relative L2 error {worst:.2e} over {len(tp)} tensors")


def test_engine_step_trains_the_tower_too():
    from rlaifv_b200.engine import DPOStepEngine
    pol, params, batch, _ = full_setup(B=2, seed=4)
    before = pol.tower.flat.clone()
    eng = DPOStepEngine(pol, lr=1e-3, total_steps=10, constant_lr=True)
    names = {b.name for b in eng.opt.buckets}
    assert {"eva_embed", "eva0", "eva2", "resampler", "embed", "layer0", "head"} <= names
    out = pol.forward_logps(batch["concatenated_input_ids"], batch["concatenated_labels"], batch["images"],
                            keep_stash=False)
    B = 2
    batch["ref_win_logp"], batch["ref_rej_logp"] = out["logp"][:B].float().cpu(), out["logp"][B:].float().cpu()
    batch["beta"] = 0.1
    m0 = float(eng.train_step(batch)[0])
    m1 = float(eng.train_step(batch)[0])
    eng.opt.wait_all()
    torch.cuda.synchronize()
    assert abs(m0 - 0.693147) < 2e-3 and m1 < m0                       # self-referenced step 0 = ln 2, then it learns
    changed = pol.tower.flat != before
    assert bool(changed.any())
    d = pol.tower.dims
    pad = pol.tower.p["b0.qkv_w"].view(3, d.num_heads, 128, d.embed_dim)[:, :, d.head_dim:]
    assert float(pad.float().abs().max()) == 0.0                       # AdamW keeps the head padding at exactly zero
    # micro-batched step (gradient accumulation through the tower) runs and stays finite
    eng.micro_pairs = 1
    m2 = float(eng.train_step(batch)[0])
    assert m2 == m2
