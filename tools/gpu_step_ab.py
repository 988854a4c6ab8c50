"""In-process A/B of step-level switches on the config-(b) DPO step (same process, same box, alternating settings —
the only comparison that is reliable on the power-capped part):
  tma_store  : GEMM epilogue through swizzled smem + cp.async.bulk.tensor stores (1) vs direct 16-byte stores (0)
  split_bwd  : attention backward as dK/dV + dQ kernels (1) vs the fused kernel with fp32 dQ atomics (0)
  fwd_variant: attention forward ping-pong kernel (1) vs single-tile kernel (0)

    python tools/gpu_step_ab.py
    python tools/gpu_step_ab.py --l2      # GEMM L2 raster / eviction-hint policy: per-shape (auto) vs the fixed raster
    python tools/gpu_step_ab.py --swiglu  # SwiGLU backward fused into the down-projection dgrad epilogue vs separate
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from rlaifv_b200 import lib
from rlaifv_b200.engine import DPOStepEngine
from rlaifv_b200.model import LlavaDims, LlavaDPOPolicy

L = lib.load()
B = bench.PAIRS_PER_GPU
pol = LlavaDPOPolicy(LlavaDims(), torch.device("cuda", 0), seed=0)
eng = DPOStepEngine(pol, lr=5e-7, total_steps=2672, micro_pairs=B)
hb = bench.synthetic_batch(0, 0, B)
out = pol.forward_logps(hb["concatenated_input_ids"], hb["concatenated_labels"], hb["images"], keep_stash=False)
hb["ref_win_logp"], hb["ref_rej_logp"] = out["logp"][:B].float().cpu(), out["logp"][B:].float().cpu()
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}


def setting(tma, split, fwd):
    L.rlaifv_gemm_set_tuning(0, 0 if tma else 8)
    pol.split_attention_bwd = bool(split)
    L.rlaifv_attention_set_variant(fwd)


def measure(n=4):
    eng.train_step(batch)
    eng.opt.wait_all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.train_step(batch)
    eng.opt.wait_all()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(2):
    eng.train_step(batch)
if "--swiglu" in sys.argv:     # SwiGLU backward in the down-projection dgrad epilogue vs the separate pass
    res = {True: [], False: []}
    for rnd in range(5):
        for v in (False, True):
            pol.fuse_swiglu_bwd = v
            res[v].append(measure())
    pol.fuse_swiglu_bwd = True
    for v, name in ((False, "separate swiglu_bwd pass"), (True, "fused into the dgrad epilogue (default)")):
        print("%-42s: %s ms/step  mean %.1f" % (name, ["%.1f" % t for t in res[v]], sum(res[v]) / len(res[v])))
    sys.exit(0)
if "--splitk" in sys.argv:     # K-sliced passes for the longest-K launches only (dgrad of gate|up, K = 22016)
    variants = ((0, 0), (2, 20000), (2, 16384))
    res = {v: [] for v in variants}
    for rnd in range(4):
        for v in variants:
            L.rlaifv_gemm_set_split_k(*v)
            res[v].append(measure(3))
    L.rlaifv_gemm_set_split_k(0, 0)
    for v in variants:
        print("split_k n=%d min_k=%d: %s ms/step  mean %.1f" % (v + (["%.1f" % t for t in res[v]], sum(res[v]) / len(res[v]))))
    sys.exit(0)
if "--l2" in sys.argv:
    res = {-1: [], 0: []}
    for rnd in range(5):
        for l2 in (0, -1):
            L.rlaifv_gemm_set_l2(l2)
            res[l2].append(measure())
    L.rlaifv_gemm_set_l2(-1)
    for l2, name in ((0, "fixed raster (16 row blocks per group, no hints)"), (-1, "per-shape policy (default)")):
        print("%-50s: %s ms/step  mean %.1f" % (name, ["%.1f" % t for t in res[l2]], sum(res[l2]) / len(res[l2])))
    sys.exit(0)
configs = [(1, 1, 1), (0, 1, 1), (1, 0, 1), (0, 0, 1), (0, 0, 0)]
res = {c: [] for c in configs}
for rnd in range(3):
    for c in configs:
        setting(*c)
        res[c].append(measure())
for c in configs:
    print("tma_store=%d split_bwd=%d fwd_variant=%d : %s ms/step" % (c + (["%.1f" % t for t in res[c]],)))
setting(1, 1, 1)
