"""CPU checkers for the batch feeder (xfm_b200/feed.py, csrc/feed.cu).  TEST INFRASTRUCTURE ONLY: imported by tests/ (and by
nothing under xfm_b200/); never on a product path.

The reference's loaders are Python over third-party image code, so the oracle for the pixel path is that code itself where it
is importable (PIL / torchvision are in the image, here and on the GPU box) plus plain restatements in torch / numpy:

  * to_tensor_normalize   transforms.ToTensor() + transforms.Normalize(mean, std) (+ hflip), dataset/__init__.py:26-35, as
                          (u8 / 255 - mean) / std in fp32 — pinned to torchvision by tests/test_feed_cpu.py
  * pil_crop_resize       PIL's own crop(box).resize(size, BICUBIC): what RandomResizedCrop / Resize with
                          InterpolationMode.BICUBIC (dataset/__init__.py:28-30,63-67) execute through torchvision
  * resample_with_taps    Pillow's ImagingResample (src/libImaging/Resample.c: horizontal pass into uint8, then vertical,
                          clip8((2^21 + sum tap * pixel) >> 22)) evaluated in numpy integer arithmetic from given tap tables
The sample / batch assembly (integer and index work) is pinned by tests/golden/feed.json, which tools/make_golden_feed.py
writes by running the reference's unmodified dataset code, and by tests/feed_probe.py (live differential)."""
import torch


def to_tensor_normalize(u8, mean, std, flip=None):
    """u8 uint8 [B, H, W, 3] -> fp32 [B, 3, H, W]; flip: optional [B] flags (1 = mirror the row, torchvision hflip)."""
    x = u8.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    x = (x - torch.tensor(mean, dtype=torch.float32)[None, :, None, None]) / torch.tensor(std, dtype=torch.float32)[None, :, None, None]
    if flip is not None:
        x = torch.where(flip.bool()[:, None, None, None], x.flip(3), x)
    return x


def pil_crop_resize(img, box, out_h, out_w):
    """img uint8 ndarray [H, W, 3]; box (x0, y0, x1, y1) or None -> uint8 ndarray [out_h, out_w, 3] computed by PIL."""
    import numpy as np
    from PIL import Image
    pil = Image.fromarray(img)
    return np.asarray((pil if box is None else pil.crop(box)).resize((out_w, out_h), Image.BICUBIC))


def resample_with_taps(img, box, hb, hk, vb, vk):
    """The two fixed-point passes of Pillow's resampler on the crop `box` of img (uint8 ndarray [H, W, 3]) with tap tables
    hb / vb int [out, 2] (first tap, count) and hk / vk int [out, K] (taps * 2^22; entries beyond the count are zero)."""
    import numpy as np
    x0, y0, x1, y1 = box
    src = img[y0:y1, x0:x1].astype(np.int64)
    hb, hk, vb, vk = (np.asarray(t).astype(np.int64) for t in (hb, hk, vb, vk))
    idx = np.minimum(hb[:, :1] + np.arange(hk.shape[1])[None, :], src.shape[1] - 1)
    tmp = np.clip(((src[:, idx, :] * hk[None, :, :, None]).sum(2) + (1 << 21)) >> 22, 0, 255)
    idx = np.minimum(vb[:, :1] + np.arange(vk.shape[1])[None, :], src.shape[0] - 1)
    return np.clip(((tmp[idx, :, :] * vk[:, :, None, None]).sum(1) + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
