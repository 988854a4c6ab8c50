"""CPU model of the GEMM's DRAM read traffic under the persistent tile schedule: an LRU cache of `cap` MB in front of
DRAM, 74 CTA pairs walking their 256x256 tiles in lock-step over K (64 per step), operands fetched as 256x64 bf16
blocks (32 KB). Calibrated against `ncu` (profiles/ncu_gemm_traffic.json: forward 0.68 GB, dgrad 2.43 GB, wgrad
2.53 GB read per launch at GROUP_M = 16) it predicts what a different rasterisation group or a split-K order would read.

    python tools/l2_raster_model.py            # table for the qkv-shaped launches of the 7B layer
    python tools/l2_raster_model.py --raster   # m- vs n-grouped walks against the measured sweep (profiles/r02u_*)
"""
import sys
from collections import OrderedDict

BLK = 256 * 64 * 2          # one operand block per k-step per tile row/col (bytes)
PAIRS = 74


def tile_order(num_m, num_n, group_m, n_grouped=False):
    if n_grouped:                       # same walk with the roles of m and n exchanged (gemm.cu: tile_coords)
        return [(m, n) for (n, m) in tile_order(num_n, num_m, group_m)]
    out = []
    per_group = group_m * num_n
    for t in range(num_m * num_n):
        g = t // per_group
        first = g * group_m
        gsz = min(num_m - first, group_m)
        r = t - g * per_group
        out.append((first + r % gsz, r // gsz))
    return out


def simulate(M, N, K, group_m, cap_mb, split_k=1, n_grouped=False):
    num_m, num_n, num_k = -(-M // 256), -(-N // 256), -(-K // 64)
    cache, cap = OrderedDict(), int(cap_mb * 2 ** 20 // BLK)
    miss = 0

    def touch(key):
        nonlocal miss
        if key in cache:
            cache.move_to_end(key)
            return
        miss += 1
        cache[key] = True
        if len(cache) > cap:
            cache.popitem(last=False)

    order = tile_order(num_m, num_n, group_m, n_grouped)
    ks = [range(s * num_k // split_k, (s + 1) * num_k // split_k) for s in range(split_k)]
    for krange in ks:                                   # split-K: all tiles for one K slice, then the next (C += ...)
        for w in range(0, len(order), PAIRS):
            wave = order[w:w + PAIRS]
            for kb in krange:
                for (m, n) in wave:
                    touch(("A", m, kb))
                    touch(("B", n, kb))
    extra_c = (split_k - 1) * 2 * M * N * 2             # re-read + re-write of C for the accumulating passes
    return miss * BLK, extra_c


def raster_table():
    """Model vs the measured sweep (ncu dram__bytes_read.sum per launch, profiles/r02u_gemm_l2_sweep_dram_bytes.txt)."""
    measured = {   # label -> (group, n_grouped): GB read
        "dgrad qkv (M=18160 N=4096 K=12288)": ((18160, 4096, 12288), {(16, 0): 3.20, (8, 0): 2.09, (4, 0): 2.47,
                                                                     (32, 0): 5.51, (8, 1): 1.80, (4, 1): 2.22}),
        "wgrad qkv (M=12288 N=4096 K=18160)": ((12288, 4096, 18160), {(16, 0): 4.32, (8, 0): 2.69, (4, 0): 2.71,
                                                                     (32, 0): 5.65, (8, 1): 2.48, (4, 1): 2.45}),
        "dgrad gate|up (M=18160 N=4096 K=22016)": ((18160, 4096, 22016), {(16, 0): 8.28, (8, 0): 5.84, (4, 0): 5.36,
                                                                         (32, 0): 11.84, (8, 1): 5.64, (4, 1): 4.78}),
        "forward qkv (M=18160 N=12288 K=4096)": ((18160, 12288, 4096), {(16, 0): 0.68, (8, 0): 1.05, (4, 0): 1.94,
                                                                       (32, 0): 1.63, (8, 1): 1.00, (4, 1): 1.88}),
    }
    caps = (48, 64, 96)
    print("DRAM read per launch, GB: measured (ncu) vs lock-step LRU model at three effective L2 capacities")
    for name, ((M, N, K), meas) in measured.items():
        print(f"\n{name}: operands {(M * K + N * K) * 2 / 1e9:.2f} GB")
        print("  raster      | measured | " + " | ".join(f"model {c:3d} MB" for c in caps))
        for (g, ng), v in meas.items():
            row = [f"{simulate(M, N, K, g, c, 1, bool(ng))[0] / 1e9:12.2f}" for c in caps]
            print(f"  {'n' if ng else 'm'}-grouped {g:2d} | {v:8.2f} | " + " | ".join(row))


def main():
    if "--raster" in sys.argv:
        return raster_table()
    shapes = {"forward  (M=18160 N=12288 K=4096)": (18160, 12288, 4096, 0.676),
              "dgrad    (M=18160 N=4096  K=12288)": (18160, 4096, 12288, 2.433),
              "wgrad    (M=12288 N=4096  K=18160)": (12288, 4096, 18160, 2.527)}
    caps = (64, 96, 126)
    print("DRAM read per launch, GB (model) — measured column = ncu at GROUP_M=16, no split-K")
    for name, (M, N, K, meas) in shapes.items():
        alg = (M * K + N * K) * 2 / 1e9
        print(f"\n{name}: algorithmic operand bytes {alg:.2f} GB, measured {meas:.2f} GB")
        print("  cap MB | " + " | ".join(f"g={g:<2d} sk={sk}" for g, sk in
                                         ((16, 1), (8, 1), (6, 1), (4, 1), (16, 2), (8, 2), (8, 4))))
        for cap in caps:
            row = []
            for g, sk in ((16, 1), (8, 1), (6, 1), (4, 1), (16, 2), (8, 2), (8, 4)):
                rd, extra = simulate(M, N, K, g, cap, sk)
                row.append(f"{(rd + extra) / 1e9:9.2f}")
            print(f"  {cap:6d} | " + " | ".join(row))


if __name__ == "__main__":
    sys.exit(main())
