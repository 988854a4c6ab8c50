"""CLIP image preprocessing for the data pipeline — what `vision_tower.image_processor` is in the reference
(`CLIPImageProcessor.from_pretrained(vision_tower_name)`, llava/model/multimodal_encoder/clip_encoder.py:29;
bound to the dataset as `lambda x: image_processor(x)['pixel_values'][0]`, muffin/train/train_llava15.py:244).

openai/clip-vit-large-patch14-336's preprocessor_config.json: RGB -> resize so the SHORT edge is 336 (bicubic)
-> center crop 336x336 -> 1/255 -> (x - mean) / std, channels first, float32.  Host-side PIL/numpy code: the image
decode + resize runs in the DataLoader workers, not on the GPU path.
"""
import json
import os

import numpy as np

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class ClipImageProcessor:
    """Callable like HF's processor: `proc(pil_image)['pixel_values'][0]` is a float32 [3, S, S] array.
    Attribute names (`crop_size`, `size`, `image_mean`, `image_std`) follow HF's, because the reference's data code
    reads them (`image_processor.crop_size` for image-less samples, `image_mean` for the 'pad' aspect mode)."""

    def __init__(self, shortest_edge=336, crop=336, image_mean=OPENAI_CLIP_MEAN, image_std=OPENAI_CLIP_STD,
                 do_center_crop=True, rescale_factor=1.0 / 255.0):
        self.size = {"shortest_edge": int(shortest_edge)}
        self.crop_size = {"height": int(crop), "width": int(crop)}
        self.image_mean = [float(x) for x in image_mean]
        self.image_std = [float(x) for x in image_std]
        self.do_center_crop = bool(do_center_crop)
        self.rescale_factor = float(rescale_factor)

    @classmethod
    def from_pretrained(cls, path, image_size=336):
        """Reads `<path>/preprocessor_config.json` when the vision tower is a local directory; otherwise (hub name,
        no network) the openai/clip-vit-large-patch14-336 defaults at `image_size`."""
        cfg_path = os.path.join(path, "preprocessor_config.json") if path and os.path.isdir(path) else None
        if cfg_path and os.path.exists(cfg_path):
            with open(cfg_path) as f:
                c = json.load(f)
            size = c.get("size", image_size)
            short = size.get("shortest_edge", image_size) if isinstance(size, dict) else int(size)
            crop = c.get("crop_size", image_size)
            crop = crop.get("height", image_size) if isinstance(crop, dict) else int(crop)
            return cls(short, crop, c.get("image_mean", OPENAI_CLIP_MEAN), c.get("image_std", OPENAI_CLIP_STD),
                       c.get("do_center_crop", True), c.get("rescale_factor", 1.0 / 255.0))
        return cls(image_size, image_size)

    def preprocess(self, image):
        from PIL import Image
        if not isinstance(image, Image.Image):
            image = Image.fromarray(np.asarray(image))
        image = image.convert("RGB")
        w, h = image.size
        short, long = (w, h) if w <= h else (h, w)
        s = self.size["shortest_edge"]
        new_short, new_long = s, int(s * long / short)
        nw, nh = (new_short, new_long) if w <= h else (new_long, new_short)
        if (nw, nh) != (w, h):
            image = image.resize((nw, nh), resample=Image.BICUBIC)
        arr = np.asarray(image, dtype=np.uint8)
        if self.do_center_crop:
            ch, cw = self.crop_size["height"], self.crop_size["width"]
            top, left = (nh - ch) // 2, (nw - cw) // 2
            if top < 0 or left < 0:          # smaller than the crop: zero-pad symmetrically (HF center_crop does too)
                pt, pl = max(0, -top), max(0, -left)
                pad = np.zeros((max(nh, ch), max(nw, cw), 3), dtype=np.uint8)
                pad[pt:pt + nh, pl:pl + nw] = arr
                arr, top, left = pad, max(0, top), max(0, left)
            arr = arr[top:top + ch, left:left + cw]
        x = arr.astype(np.float32) * np.float32(self.rescale_factor)
        x = (x - np.asarray(self.image_mean, dtype=np.float32)) / np.asarray(self.image_std, dtype=np.float32)
        return np.ascontiguousarray(x.transpose(2, 0, 1))

    def __call__(self, images, return_tensors=None):
        many = isinstance(images, (list, tuple))
        out = [self.preprocess(im) for im in (images if many else [images])]
        if return_tensors == "pt":
            import torch
            return {"pixel_values": torch.from_numpy(np.stack(out))}
        return {"pixel_values": out}


class PixelValues:
    """`lambda x: image_processor(x)['pixel_values'][0]` (muffin/train/train_llava15.py:244) as an object, so the
    data code can still read `crop_size` / `image_mean` from it; returns a float32 torch tensor [3, S, S]."""

    def __init__(self, processor):
        self.processor = processor
        self.crop_size = processor.crop_size
        self.image_mean = processor.image_mean

    def __call__(self, image):
        import torch
        return torch.from_numpy(self.processor(image)["pixel_values"][0])


class SquareResizeProcessor:
    """OmniLMM's evaluation transform (`build_transform(is_train=False, input_size, std_mode='OPENAI_CLIP')`,
    omnilmm/model/utils.py:455-460): resize to input_size x input_size (bicubic, aspect ratio NOT kept) -> [0,1] ->
    CLIP mean/std. The training transform of the reference is the same resize preceded by a RandomResizedCrop of
    scale (0.9999, 1) — i.e. the whole image — so the deterministic form is used for both."""

    def __init__(self, input_size=448, image_mean=OPENAI_CLIP_MEAN, image_std=OPENAI_CLIP_STD):
        self.input_size = int(input_size)
        self.crop_size = {"height": self.input_size, "width": self.input_size}
        self.image_mean = [float(x) for x in image_mean]
        self.image_std = [float(x) for x in image_std]

    def __call__(self, image):
        import torch
        from PIL import Image
        if not isinstance(image, Image.Image):
            image = Image.fromarray(np.asarray(image))
        image = image.convert("RGB").resize((self.input_size, self.input_size), resample=Image.BICUBIC)
        x = np.asarray(image, dtype=np.float32) / np.float32(255.0)
        x = (x - np.asarray(self.image_mean, dtype=np.float32)) / np.asarray(self.image_std, dtype=np.float32)
        return torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1)))
