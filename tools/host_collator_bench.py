"""Host-side cost of the preference collator (`DataCollatorForDPODataset.__call__`,
muffin/train/train_muffin.py:43-112 — padding + difflib token diff per pair) at the config-(b) shape: 8 pairs of
560-token sequences (48-token prompt + 512-token responses that share a prefix, as real chosen / rejected pairs do).

    python tools/host_collator_bench.py            # CPU only

Prints batches/s and pairs/s of ONE worker process, next to the GPU step's appetite (11.4 pairs/s per GPU), i.e. how many
DataLoader workers per GPU keep the engine fed (the shipped recipe uses 16, script/train/llava15_train.sh:22)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlaifv_b200.collator import DataCollatorForDPODataset


class Tok:
    pad_token_id = 0


def make_instances(B, share, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(B):
        prompt = torch.randint(3, 32000, (48,), generator=g)
        prompt[0], prompt[35] = 1, -200
        base = torch.randint(3, 32000, (512,), generator=g)
        win, rej = base.clone(), base.clone()
        n_edit = int(512 * (1 - share))
        idx = torch.randperm(512, generator=g)[:n_edit]
        rej[idx] = torch.randint(3, 32000, (n_edit,), generator=g)

        def one(resp, kind):
            ids = torch.cat([prompt, resp])
            return {"input_ids": ids, "labels": torch.cat([torch.full((48,), -100), resp]), "image": torch.zeros(3, 336, 336),
                    f"ref_{kind}_logp": -500.0, f"ref_{kind}_avg_logp": -1.0, f"ref_{kind}_per_token_logp": [0.0] * 1200}
        out.append((one(rej, "rej"), one(win, "win")))
    return out


def main():
    torch.set_num_threads(1)
    coll = DataCollatorForDPODataset(tokenizer=Tok(), beta=0.1, mod_token_weight=1.0)
    for share in (0.9, 0.5, 0.0):
        inst = make_instances(8, share, 0)
        coll(inst)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 3.0:
            coll(inst)
            n += 1
        dt = (time.perf_counter() - t0) / n
        print("responses share %3.0f%% of their tokens: %.1f ms per 8-pair batch = %.0f pairs/s per worker "
              "(GPU step consumes ~11.4 pairs/s per GPU => %.2f workers per GPU)"
              % (100 * share, dt * 1e3, 8 / dt, 11.4 / (8 / dt)))


def image_path_cost():
    """JPEG decode + CLIP preprocessing (resize shortest edge 336 bicubic, center crop, normalise) per image — the other
    per-sample host cost (muffin/data/datasets.py:92 bytes_to_PIL_image + muffin/train/train_llava15.py:244)."""
    import io
    import numpy as np
    from PIL import Image
    from rlaifv_b200.image_processing import ClipImageProcessor, PixelValues
    proc = PixelValues(ClipImageProcessor(336, 336))
    rng = np.random.RandomState(0)
    base = rng.randint(0, 256, (60, 80, 3), dtype=np.uint8)
    img = Image.fromarray(base).resize((640, 480), Image.BICUBIC)          # a photo-like (smooth) 640x480 image
    buf = io.BytesIO()
    img.save(buf, format="JPEG", quality=90)
    raw = buf.getvalue()
    proc(Image.open(io.BytesIO(raw)).convert("RGB"))
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 3.0:
        proc(Image.open(io.BytesIO(raw)).convert("RGB"))
        n += 1
    dt = (time.perf_counter() - t0) / n
    print("JPEG (640x480, %d KB) decode + CLIP preprocess to 336x336: %.1f ms per image = %.0f pairs/s per worker "
          "(one image per pair) => %.2f workers per GPU" % (len(raw) // 1024, dt * 1e3, 1 / dt, 11.4 * dt))


if __name__ == "__main__":
    main()
    image_path_cost()
