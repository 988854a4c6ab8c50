"""In-process A/B of the experimental split-K order (rlaifv_gemm_set_split_k) on the 7B layer's long-K launches, and —
with --step — on whole DPO steps (same process, same box, alternating settings; only this kind of comparison is
reliable on the power-capped part). Run on a B200:

    python tools/gpu_gemm_splitk_ab.py            # dgrad / wgrad shapes, n = 1, 2, 4
    python tools/gpu_gemm_splitk_ab.py --step     # 4 training steps per setting, twice
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlaifv_b200 import lib, ops

L = lib.load()
dev = "cuda"


def bench_shape(name, make, reps=20):
    a, b, out, kw = make()
    res = {}
    for rnd in range(2):
        for n in (1, 2, 4):
            L.rlaifv_gemm_set_split_k(n, 8192)
            for _ in range(3):
                ops.gemm(a, b, out, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ops.gemm(a, b, out, **kw)
            e1.record()
            torch.cuda.synchronize()
            res.setdefault(n, []).append(e0.elapsed_time(e1) / reps)
    L.rlaifv_gemm_set_split_k(0, 0)
    print(name, {n: ["%.3f ms" % t for t in ts] for n, ts in res.items()})


def shapes():
    M, H, F = 18160, 4096, 11008
    x = torch.randn(M, H, device=dev).bfloat16()
    dy3 = torch.randn(M, 3 * H, device=dev).bfloat16()
    w3 = (torch.randn(3 * H, H, device=dev) * 0.02).bfloat16()
    dgu = torch.randn(M, 2 * F, device=dev).bfloat16()
    wgu = (torch.randn(2 * F, H, device=dev) * 0.02).bfloat16()
    act = torch.randn(M, F, device=dev).bfloat16()
    dx = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
    yield "dgrad qkv   K=12288", lambda: (dy3, w3, dx, dict(b_mn=True))
    yield "dgrad gu    K=22016", lambda: (dgu, wgu, dx, dict(b_mn=True))
    yield "wgrad qkv   K=18160", lambda: (dy3, x, torch.empty(3 * H, H, device=dev, dtype=torch.bfloat16),
                                          dict(a_mn=True, b_mn=True))
    yield "wgrad gu    K=18160", lambda: (dgu, x, torch.empty(2 * F, H, device=dev, dtype=torch.bfloat16),
                                          dict(a_mn=True, b_mn=True))
    yield "wgrad down  K=18160", lambda: (dx, act, torch.empty(H, F, device=dev, dtype=torch.bfloat16),
                                          dict(a_mn=True, b_mn=True))


def step_ab():
    import bench
    from rlaifv_b200.engine import DPOStepEngine
    from rlaifv_b200.model import LlavaDims, LlavaDPOPolicy
    B = bench.PAIRS_PER_GPU
    pol = LlavaDPOPolicy(LlavaDims(), torch.device("cuda", 0), seed=0)
    eng = DPOStepEngine(pol, lr=5e-7, total_steps=2672, micro_pairs=B)
    hb = bench.synthetic_batch(0, 0, B)
    out = pol.forward_logps(hb["concatenated_input_ids"], hb["concatenated_labels"], hb["images"], keep_stash=False)
    hb["ref_win_logp"], hb["ref_rej_logp"] = out["logp"][:B].float().cpu(), out["logp"][B:].float().cpu()
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
    for _ in range(2):
        eng.train_step(batch)
    for rnd in range(2):
        for n in (1, 2, 4):
            L.rlaifv_gemm_set_split_k(n, 8192)
            eng.train_step(batch)
            eng.opt.wait_all()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                m = eng.train_step(batch)
            eng.opt.wait_all()
            e1.record()
            torch.cuda.synchronize()
            print("round %d split_k=%d: %.1f ms/step, loss %.5f" % (rnd, n, e0.elapsed_time(e1) / 4, float(m[0])))
    L.rlaifv_gemm_set_split_k(0, 0)


if __name__ == "__main__":
    if "--step" in sys.argv:
        step_ab()
    else:
        for name, mk in shapes():
            bench_shape(name, mk)
