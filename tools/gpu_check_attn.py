"""GPU check of the tcgen05 attention forward/backward against fp32 torch attention."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlaifv_b200 import ops

torch.manual_seed(0)
dev = "cuda"
from rlaifv_b200 import lib as _lib
_variant = int(os.environ.get("RLAIFV_ATT_VARIANT", "1"))
_lib.load().rlaifv_attention_set_variant(_variant)
print("attention forward variant", _variant, flush=True)
fails = 0

def check(name, got, ref, tol):
    global fails
    got = got.float(); ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    bad = not (err <= tol * scale) or not torch.isfinite(got).all().item()
    print(f"{'FAIL' if bad else 'ok  '} {name}: max_abs_err={err:.4g} ref_max={scale:.4g}", flush=True)
    fails += int(bad)

def ref_attn(q, k, v, causal, scale):
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        S = q.shape[-2]
        mask = torch.ones(S, S, device=q.device, dtype=torch.bool).tril()
        s = s.masked_fill(~mask, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.softmax(s, dim=-1)
    return p @ v, lse

def run(nseq, S, nh, D, causal, bwd=False, qscale=1.0):
    H = nh * D
    qkv = (torch.randn(nseq * S, 3 * H, device=dev) * qscale).bfloat16()
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    scale = 1.0 / math.sqrt(D)
    out, lse = ops.attention_fwd(q, k, v, nseq, S, nh, D, causal, scale)
    torch.cuda.synchronize()
    def split(t):
        return t.float().reshape(nseq, S, nh, D).permute(0, 2, 1, 3).contiguous()
    qf, kf, vf = split(q).requires_grad_(), split(k).requires_grad_(), split(v).requires_grad_()
    ro, rl = ref_attn(qf, kf, vf, causal, scale)
    tag = f"nseq={nseq} S={S} nh={nh} D={D} causal={causal} qs={qscale}"
    check("attn_fwd out " + tag, split(out), ro, 2e-2)
    check("attn_fwd lse " + tag, lse, rl, 2e-3)
    if bwd:
        d_out = torch.randn(nseq * S, H, device=dev).bfloat16()
        ro.backward(split(d_out))
        if causal:                                    # the fused round-1 kernel's entry point is causal-only
            dq32 = torch.zeros(nseq * S, H, device=dev, dtype=torch.float32)
            dqkv = torch.zeros(nseq * S, 3 * H, device=dev, dtype=torch.bfloat16)
            ops.attention_bwd(q, k, v, out, d_out, lse, nseq, S, nh, D, scale, dq32, dqkv[:, H:2 * H], dqkv[:, 2 * H:])
            torch.cuda.synchronize()
            check("attn_bwd dq " + tag, split(dq32), qf.grad, 3e-2)
            check("attn_bwd dk " + tag, split(dqkv[:, H:2 * H]), kf.grad, 3e-2)
            check("attn_bwd dv " + tag, split(dqkv[:, 2 * H:]), vf.grad, 3e-2)
        # split backward (dK/dV kernel + dQ kernel), dq as bf16 in the q block of a fused buffer (poisoned first)
        d2 = torch.full((nseq * S, 3 * H), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.attention_bwd_split(q, k, v, out, d_out, lse, nseq, S, S, nh, D, causal, scale, d2[:, :H], d2[:, H:2 * H],
                                d2[:, 2 * H:])
        torch.cuda.synchronize()
        check("attn_bwd_split dq " + tag, split(d2[:, :H]), qf.grad, 3e-2)
        check("attn_bwd_split dk " + tag, split(d2[:, H:2 * H]), kf.grad, 3e-2)
        check("attn_bwd_split dv " + tag, split(d2[:, 2 * H:]), vf.grad, 3e-2)

cases = [
    (1, 128, 1, 128, True, False, 1.0), (1, 128, 1, 128, False, False, 1.0), (1, 256, 2, 128, True, False, 1.0),
    (2, 300, 2, 128, True, False, 1.0), (2, 1135, 4, 128, True, False, 1.0), (2, 1135, 4, 128, True, False, 4.0),
    (1, 128, 1, 64, False, False, 1.0), (2, 577, 4, 64, False, False, 1.0), (3, 5, 2, 64, False, False, 1.0),
    (2, 1025, 2, 128, False, False, 1.0), (1, 1135, 2, 128, True, False, 8.0), (2, 384, 2, 128, True, False, 1.0),
    (1, 128, 1, 128, True, True, 1.0), (1, 256, 2, 128, True, True, 1.0), (2, 300, 2, 128, True, True, 1.0),
    (2, 1135, 4, 128, True, True, 1.0), (2, 687, 4, 128, True, True, 3.0), (2, 1025, 2, 128, False, True, 1.0),
    (3, 70, 2, 128, False, True, 1.0), (1, 64, 1, 128, True, True, 1.0),
    # large logits (lazy-rescale path taken on most tiles) and a long non-causal row
    (1, 1135, 2, 128, True, False, 16.0), (2, 2049, 2, 128, False, False, 6.0), (2, 1100, 2, 64, False, False, 10.0),
]
for cse in cases:
    try:
        run(*cse[:5], bwd=cse[5], qscale=cse[6])
    except Exception as e:
        print("EXCEPTION", cse, repr(e), flush=True)
        fails += 1
        break

def bench(nseq, S, nh, D, causal, bwd, split=False):
    H = nh * D
    qkv = torch.randn(nseq * S, 3 * H, device=dev).bfloat16()
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    scale = 1.0 / math.sqrt(D)
    out, lse = ops.attention_fwd(q, k, v, nseq, S, nh, D, causal, scale)
    d_out = torch.randn(nseq * S, H, device=dev).bfloat16()
    dq32 = torch.zeros(nseq * S, H, device=dev, dtype=torch.float32)
    dqkv = torch.zeros(nseq * S, 3 * H, device=dev, dtype=torch.bfloat16)
    def f():
        if bwd and split:
            ops.attention_bwd_split(q, k, v, out, d_out, lse, nseq, S, S, nh, D, causal, scale, dqkv[:, :H],
                                    dqkv[:, H:2 * H], dqkv[:, 2 * H:])
        elif bwd:
            ops.attention_bwd(q, k, v, out, d_out, lse, nseq, S, nh, D, scale, dq32, dqkv[:, H:2 * H], dqkv[:, 2 * H:])
        else:
            ops.attention_fwd(q, k, v, nseq, S, nh, D, causal, scale, out, lse)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * nseq * nh * S * S * D * (0.5 if causal else 1.0) * (2.5 if bwd else 1.0)
    print(f"perf attn {('bwd_split' if split else 'bwd') if bwd else 'fwd'} nseq={nseq} S={S} nh={nh} D={D} causal={causal}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s (algorithmic)", flush=True)

if fails == 0:
    try:
        for var in (0, 1, 0, 1):                        # in-process A/B: single-tile kernel vs two-tile ping-pong
            _lib.load().rlaifv_attention_set_variant(var)
            print("forward variant", var, flush=True)
            bench(16, 1135, 32, 128, True, False)
            bench(4, 1025, 16, 128, False, False)
            bench(16, 577, 16, 64, False, False)
        _lib.load().rlaifv_attention_set_variant(_variant)
        bench(16, 1135, 32, 128, True, True)
        bench(16, 1135, 32, 128, True, True, split=True)
        bench(16, 577, 16, 64, False, False)
    except Exception as e:
        print("EXCEPTION", repr(e), flush=True); fails += 1
print("FAILS", fails)
sys.exit(1 if fails else 0)
