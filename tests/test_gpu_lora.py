"""-m gpu: LoRA-DPO path (BASELINE config e) — adapters on q,k,v,o,gate,up,down, base frozen, projector
trainable — against the oracle's restatement of peft's `W x + (alpha/r) B(A(x))` (dropout 0) and against fixtures of the
UNMODIFIED reference model run with merged weights W' = W + s B A (tests/golden/lora/, oracle/gen_golden_lora.py:
log-probs, losses and — via dA = s B^T dW', dB = s dW' A^T — every adapter gradient)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import llava_dpo_oracle as O

pytestmark = pytest.mark.gpu
R, ALPHA = 8, 2.0   # scaling 0.25 like the shipped r=64 / alpha=16


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def setup():
    from rlaifv_b200.model import LlavaDims, LlavaDPOPolicy
    c = O.TINY
    dims = LlavaDims(vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                     num_layers=c.num_layers, num_heads=c.num_heads, clip_hidden=c.clip_hidden,
                     clip_intermediate=c.clip_intermediate, clip_layers=c.clip_layers, clip_heads=c.clip_heads,
                     image_size=c.image_size, patch_size=c.patch_size)
    params = O.make_params(c, seed=0, scale=0.4)
    lora = O.make_lora_params(c, r=R, seed=3)
    pol = LlavaDPOPolicy(dims, "cuda", hf_state=params)
    store = pol.enable_lora(r=R, alpha=ALPHA)
    store.load_hf(lora)
    return pol, params, lora


def test_lora_forward_backward_match_oracle():
    from rlaifv_b200 import ops
    pol, params, lora = setup()
    c = O.TINY
    batch = O.synthetic_pair_batch(c, 2, 24, 30, seed=31, image_pos=7, ragged=True)
    ids, labels, images = batch["concatenated_input_ids"], batch["concatenated_labels"], batch["images"]
    out = pol.forward_logps(ids, labels, images, keep_stash=True)
    # oracle (bf16 op order) with adapters; grads in fp32 for the gradient check
    pb = {k: v.to(torch.bfloat16) for k, v in {**params, **lora}.items()}
    ob = O.policy_logps(pb, c, ids, labels, images.to(torch.bfloat16))
    assert rel(out["logp"], ob["logp"]) <= 1e-3
    base = O.policy_logps({k: v.to(torch.bfloat16) for k, v in params.items()}, c, ids, labels, images.to(torch.bfloat16))
    assert rel(ob["logp"], base["logp"]) > 1e-3           # the adapters really change the output
    pf = {k: v.clone().float().requires_grad_(True) for k, v in {**params, **lora}.items()}
    of = O.policy_logps(pf, c, ids, labels, images)
    B = 2
    rw = torch.tensor([-150.0, -160.0])
    rr = torch.tensor([-151.0, -158.0])
    losses, _, _ = O.dpo_loss(of["logp"][:B], of["logp"][B:], rw, rr, 0.1)
    losses.mean().backward()
    _, _, _, dpw, dpr, out9 = ops.dpo_loss(out["logp"][:B].contiguous(), out["logp"][B:].contiguous(), rw.cuda(), rr.cuda(), 0.1)
    pol.store.grad.zero_()
    pol.backward_logps(torch.cat([dpw, dpr]).contiguous())
    torch.cuda.synchronize()
    assert abs(float(out9[0]) - float(losses.mean())) <= 2e-2 * max(1.0, abs(float(losses.mean())))
    gl = pol.lora.hf_views(grads=True)
    worst = 0.0
    for name in ("model.layers.0.self_attn.q_proj.lora_A.weight", "model.layers.0.self_attn.v_proj.lora_B.weight",
                 "model.layers.1.self_attn.o_proj.lora_A.weight", "model.layers.1.self_attn.o_proj.lora_B.weight",
                 "model.layers.0.mlp.gate_proj.lora_B.weight", "model.layers.1.mlp.up_proj.lora_A.weight",
                 "model.layers.0.mlp.down_proj.lora_A.weight", "model.layers.1.mlp.down_proj.lora_B.weight",
                 "model.layers.1.self_attn.k_proj.lora_A.weight"):
        ref = pf[name].grad
        got = gl[name].float().cpu()
        nr = float((got.double().norm() - ref.double().norm()).abs() / (ref.double().norm() + 1e-30))
        er = rel(got, ref)
        print(f"{name}: rel max err {er:.3e}, norm diff {nr:.3e}")
        worst = max(worst, er)
        assert nr <= 5e-2 and er <= 1e-1, name
    # projector is trainable, everything else in the base store stays untouched (frozen)
    g = pol.store.hf_grad_views()
    assert rel(g["model.mm_projector.2.weight"].float(), pf["model.mm_projector.2.weight"].grad) <= 1e-1
    for name in ("lm_head.weight", "model.embed_tokens.weight", "model.layers.0.self_attn.q_proj.weight",
                 "model.layers.1.mlp.down_proj.weight", "model.norm.weight"):
        assert float(g[name].float().abs().max()) == 0.0, name


def test_lora_engine_step_updates_only_adapters_and_projector():
    from rlaifv_b200.engine import DPOStepEngine
    pol, params, lora = setup()
    before_base = pol.store.flat.clone()
    before_lora = pol.lora.flat.clone()
    eng = DPOStepEngine(pol, lr=1e-3, total_steps=10, constant_lr=True)
    assert {b.name for b in eng.opt.buckets} == {f"lora{i}" for i in range(O.TINY.num_layers)} | {"projector"}
    batch = O.synthetic_pair_batch(O.TINY, 2, 24, 20, seed=5, image_pos=7)
    batch["ref_win_logp"] = torch.tensor([-100.0, -101.0])
    batch["ref_rej_logp"] = torch.tensor([-100.5, -100.0])
    batch["beta"] = 0.1
    m = eng.train_step(batch)
    torch.cuda.synchronize()
    assert float(m[0]) == float(m[0])
    assert not torch.equal(pol.lora.flat, before_lora)
    proj = next(b for b in pol.store.buckets if b.name == "projector")
    changed = pol.store.flat != before_base
    assert bool(changed[proj.start:proj.start + proj.size].any())
    changed[proj.start:proj.start + proj.size] = False
    assert not bool(changed.any())                      # base weights, embeddings, norms, lm_head frozen


def test_lora_dropout_kernels():
    from rlaifv_b200 import ops
    x = torch.randn(1 << 16, 64, device="cuda").to(torch.bfloat16)
    p = 0.05
    a = ops.dropout_fwd(x, p, seed=1234)
    b = ops.dropout_fwd(x, p, seed=1234)
    c = ops.dropout_fwd(x, p, seed=1235)
    assert torch.equal(a, b) and not torch.equal(a, c)                     # deterministic per seed
    dropped = (a == 0) & (x != 0)
    frac = float(dropped.float().mean())
    assert abs(frac - p) < 2e-3                                            # 4M samples: sigma ~ 1e-4
    kept = ~dropped
    assert torch.equal(a[kept], (x.float()[kept] / (1 - p)).to(torch.bfloat16))
    g = torch.randn_like(x)
    dx = torch.zeros_like(x)
    ops.dropout_bwd_add(dx, g, p, seed=1234)
    assert torch.equal(dx != 0, kept & (g != 0))                           # same mask as the forward
    assert torch.equal(dx[kept], (g.float()[kept] / (1 - p)).to(torch.bfloat16))


def test_lora_with_dropout_matches_oracle_given_the_same_masks():
    """p = 0.25: the CUDA path regenerates its masks from seeds; feed exactly those masks to the oracle."""
    from rlaifv_b200 import ops
    pol, params, lora = setup()
    pol.lora.dropout = 0.25
    c = O.TINY
    batch = O.synthetic_pair_batch(c, 2, 24, 30, seed=41, image_pos=7, ragged=True)
    ids, labels, images = batch["concatenated_input_ids"], batch["concatenated_labels"], batch["images"]
    out = pol.forward_logps(ids, labels, images, keep_stash=True)
    nseq, T = ids.shape[0], out["T"]
    H, F_ = c.hidden_size, c.intermediate_size
    masks = {}
    for i, ls in enumerate(pol._stash["layers"]):
        for group, width in (("qkv", H), ("o", H), ("gu", H), ("down", F_)):
            ones = torch.ones(nseq * T, width, device="cuda", dtype=torch.bfloat16)
            masks[f"{i}.{group}"] = ops.dropout_fwd(ones, 0.25, ls["seed_" + group]).float().cpu().reshape(nseq, T, width)
    assert 0.2 < float((masks["0.qkv"] == 0).float().mean()) < 0.3
    pf = {k: v.clone().float().requires_grad_(True) for k, v in {**params, **lora}.items()}
    of = O.policy_logps(pf, c, ids, labels, images, lora_drop=masks)
    assert rel(out["logp"], of["logp"].detach()) <= 2e-3
    nodrop = O.policy_logps({k: v.detach() for k, v in pf.items()}, c, ids, labels, images)
    assert rel(of["logp"].detach(), nodrop["logp"]) > 1e-4                 # the masks matter
    g = torch.tensor([0.02, -0.01, -0.02, 0.01])
    (of["logp"] * g).sum().backward()
    pol.backward_logps(g.cuda().contiguous())
    torch.cuda.synchronize()
    gl = pol.lora.hf_views(grads=True)
    for name in ("model.layers.0.self_attn.q_proj.lora_A.weight", "model.layers.1.self_attn.o_proj.lora_B.weight",
                 "model.layers.0.mlp.up_proj.lora_A.weight", "model.layers.1.mlp.down_proj.lora_A.weight",
                 "model.layers.0.mlp.down_proj.lora_B.weight"):
        ref, got = pf[name].grad, gl[name].float().cpu()
        nr = float((got.double().norm() - ref.double().norm()).abs() / (ref.double().norm() + 1e-30))
        print(f"{name}: rel max err {rel(got, ref):.3e}, norm diff {nr:.3e}")
        assert nr <= 5e-2 and rel(got, ref) <= 1e-1, name
    assert rel(pol.store.hf_grad_views()["model.mm_projector.0.weight"].float(), pf["model.mm_projector.0.weight"].grad) <= 1e-1


def test_lora_trainer_checkpoint_layout_and_resume(tmp_path):
    """LoRA run through the trainer: checkpoints hold the peft-layout adapter + non_lora_trainables.bin + config.json
    (muffin/train/train_llava15_lora.py:184-197; what llava/model/builder.py:52-86 loads), the frozen base is never
    rewritten, and resuming restores adapters, projector and optimizer state exactly."""
    import json
    import os
    from test_gpu_trainer_compat import Tok, dims, instances, make_args
    from rlaifv_b200.collator import DataCollatorForDPODataset
    from rlaifv_b200.llava_model import LlavaLlamaForCausalLM
    from rlaifv_b200.trainers import LLaVA15DPOTrainer
    params = O.make_params(O.TINY, seed=0, scale=0.4)
    data = instances(8, seed=11)
    coll = DataCollatorForDPODataset(tokenizer=Tok(), beta=0.1, mod_token_weight=1.0)

    def run(max_steps, resume):
        model = LlavaLlamaForCausalLM(dims(), "cuda", hf_state=params)
        model.policy.enable_lora(r=R, alpha=ALPHA, dropout=0.0)
        tr = LLaVA15DPOTrainer(model=model, tokenizer=Tok(), args=make_args(tmp_path, max_steps=max_steps),
                               train_dataset=data, data_collator=coll)
        tr.train(resume_from_checkpoint=resume)
        torch.cuda.synchronize()
        return model, tr

    m_full, _ = run(4, False)                                   # uninterrupted 4 steps (also writes checkpoint-2/-4)
    ck = tmp_path / "checkpoint-2"
    assert sorted(os.listdir(ck)) == ["adapter_config.json", "adapter_model.bin", "config.json",
                                      "non_lora_trainables.bin", "optimizer_rank0.pt", "trainer_state.json"]
    cfg = json.load(open(ck / "adapter_config.json"))
    assert cfg["r"] == R and cfg["lora_alpha"] == ALPHA and cfg["peft_type"] == "LORA"
    ad = torch.load(ck / "adapter_model.bin")
    assert "base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight" in ad
    assert ad["base_model.model.model.layers.1.mlp.down_proj.lora_B.weight"].shape == (O.TINY.hidden_size, R)
    assert set(torch.load(ck / "non_lora_trainables.bin")) == {
        "base_model.model.model.mm_projector.%d.%s" % (i, n) for i in (0, 2) for n in ("weight", "bias")}
    assert json.load(open(ck / "config.json"))["model_type"] == "llava_llama"
    want = m_full.policy.lora.flat.clone()
    # drop checkpoint-4 so the second run resumes from step 2 and must land on the same adapters
    for fn in os.listdir(tmp_path / "checkpoint-4"):
        os.remove(tmp_path / "checkpoint-4" / fn)
    os.rmdir(tmp_path / "checkpoint-4")
    m_res, t_res = run(4, True)
    assert t_res.state["global_step"] == 4
    # same data order, same optimizer state => same adapters (fp32 atomics in the backward may flip a last bit)
    assert float((m_res.policy.lora.flat == want).float().mean()) > 0.99
    assert float((m_res.policy.store.flat == m_full.policy.store.flat).float().mean()) > 0.999


LORA_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lora", "*.npz")))


@pytest.mark.parametrize("path", LORA_GOLDEN, ids=[os.path.basename(p) for p in LORA_GOLDEN])
def test_lora_matches_merged_weight_reference_fixture(path):
    """CUDA LoRA path vs the reference fixture: summed log-probs 1e-3 (north_star), losses, rewards, gradient norms of
    all 28 adapter matrices + the projector within 3 %, gradient samples within 6 % of the tensor's max."""
    from rlaifv_b200 import ops
    from rlaifv_b200.model import LlavaDims, LlavaDPOPolicy
    fx = np.load(path)
    c = O.TINY
    dims = LlavaDims(vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                     num_layers=c.num_layers, num_heads=c.num_heads, clip_hidden=c.clip_hidden,
                     clip_intermediate=c.clip_intermediate, clip_layers=c.clip_layers, clip_heads=c.clip_heads,
                     image_size=c.image_size, patch_size=c.patch_size)
    params = O.make_params(c, seed=0, scale=float(fx["param_scale"]))
    lora = O.make_lora_params(c, r=int(fx["r"]), seed=int(fx["lora_seed"]), b_std=float(fx["lora_b_std"]))
    assert abs(O.params_checksum(lora) - float(fx["lora_checksum"])) <= 1e-9 * float(fx["lora_checksum"])
    pol = LlavaDPOPolicy(dims, "cuda", hf_state=params)
    pol.enable_lora(r=int(fx["r"]), alpha=float(fx["alpha"])).load_hf(lora)
    ids, labels = torch.from_numpy(fx["concatenated_input_ids"]), torch.from_numpy(fx["concatenated_labels"])
    out = pol.forward_logps(ids, labels, torch.from_numpy(fx["images"]), keep_stash=True)
    B = int(fx["B"])
    assert torch.equal(out["labels"].cpu(), torch.from_numpy(fx["spliced_labels"]))
    ref = torch.cat([torch.from_numpy(fx["policy_win_logp"]), torch.from_numpy(fx["policy_rej_logp"])])
    e_sum = rel(out["logp"], ref)
    losses, cr, rj, dpw, dpr, out9 = ops.dpo_loss(out["logp"][:B].contiguous(), out["logp"][B:].contiguous(),
                                                  torch.from_numpy(fx["ref_win_logp"]).cuda(),
                                                  torch.from_numpy(fx["ref_rej_logp"]).cuda(), float(fx["beta"]))
    pol.store.grad.zero_()
    pol.lora.grad.zero_()
    pol.backward_logps(torch.cat([dpw, dpr]).contiguous())
    torch.cuda.synchronize()
    e_loss = rel(losses, fx["losses"])
    print(f"summed logp rel err vs merged-weight reference {e_sum:.2e}; losses {e_loss:.2e}")
    assert e_sum <= 1e-3
    assert e_loss <= 5e-3
    assert rel(cr, fx["chosen_rewards"]) <= 5e-3
    gl = dict(pol.lora.hf_views(grads=True))
    gl.update({k: v for k, v in pol.store.hf_grad_views().items() if "mm_projector" in k})
    worst_n = worst_s = 0.0
    for key in fx.files:
        if not key.startswith("gradsample:"):
            continue
        name = key.split(":", 1)[1]
        g = gl[name].float().flatten().cpu()
        idx = torch.linspace(0, g.numel() - 1, 64).long()
        refs = torch.from_numpy(fx[key])
        gn = float(fx["gradnorm:" + name])
        en = abs(float(g.double().norm()) - gn) / gn
        es = float((g[idx] - refs).abs().max()) / (float(refs.abs().max()) + 1e-12)
        worst_n, worst_s = max(worst_n, en), max(worst_s, es)
        assert en <= 3e-2, (name, en)
        assert es <= 6e-2, (name, es)
    print(f"adapter + projector gradients: worst norm err {worst_n:.2e}, worst sample err {worst_s:.2e}")
