"""Device-timed figures for the batch feeder (csrc/feed.cu, xfm_b200/feed.py): the ToTensor + Normalize kernel against the HBM
roofline (algorithmic bytes = 3 B read + 12 B written per pixel) and one batch end to end from pageable / pinned host memory:
uint8 crops + GPU transform vs the fp32 tensors the reference's workers produce.  Writes gpurun_out/feed.jsonl."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xfm_b200 import feed, lib  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def resize_rows():
    """crop + bicubic resize: device time of the two integer passes for a ragged batch (tables and pixels already on the
    device), the whole `feed.crop_resize` call from host tensors (tap tables on the host + H2D + kernels), and PIL on one host
    core for the same crops."""
    import time
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(0)
    rows = []
    for B, res in [(96, 224), (32, 384)]:
        images, boxes = [], []
        for _ in range(B):
            H, W = int(rng.integers(360, 641)), int(rng.integers(480, 641))
            area = H * W * rng.uniform(0.2, 1.0)                       # RandomResizedCrop(scale=(0.2, 1.0))
            ar = np.exp(rng.uniform(np.log(3 / 4), np.log(4 / 3)))
            w, h = min(W, int(round(np.sqrt(area * ar)))), min(H, int(round(np.sqrt(area / ar))))
            x0, y0 = int(rng.integers(0, W - w + 1)), int(rng.integers(0, H - h + 1))
            images.append(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8))
            boxes.append((x0, y0, x0 + w, y0 + h))
        tens = [torch.from_numpy(im) for im in images]
        plan = feed.crop_resize_plan([im.shape[:2] for im in images], boxes, res, res)
        dev = {k: plan[k].cuda() for k in ("desc", "hb", "hk", "vb", "vk")}
        packed = torch.cat([t.reshape(-1) for t in tens]).cuda()
        tmp = torch.empty(plan["tmp_bytes"], dtype=torch.uint8, device="cuda")
        out = torch.empty((B, res, res, 3), dtype=torch.uint8, device="cuda")
        ms_k = timed(lambda: lib.resize_bicubic_u8(packed, dev["desc"], dev["hb"], dev["hk"], dev["vb"], dev["vk"], tmp, out,
                                                   plan["max_rows"]))
        t0 = time.perf_counter()
        for _ in range(3):
            feed.crop_resize(tens, boxes, res, res)
        torch.cuda.synchronize()
        ms_call = (time.perf_counter() - t0) / 3 * 1e3
        t0 = time.perf_counter()
        for im, bx in zip(images, boxes):
            Image.fromarray(im).crop(bx).resize((res, res), Image.BICUBIC)
        ms_pil = (time.perf_counter() - t0) * 1e3
        crop_bytes = sum((b[2] - b[0]) * (b[3] - b[1]) * 3 for b in boxes)
        rows.append(dict(what="crop_resize", B=B, res=res, kernels_us=ms_k * 1e3, call_ms_host_to_device=ms_call,
                         pil_one_core_ms=ms_pil, src_bytes=int(plan["src_bytes"]), crop_bytes=int(crop_bytes),
                         tmp_bytes=int(plan["tmp_bytes"]), out_bytes=B * res * res * 3,
                         algorithmic_gbps=(crop_bytes + 2 * plan["tmp_bytes"] + B * res * res * 3) / ms_k * 1e-6))
    return rows


def main():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6540.0))
    rows = []
    for B, R in [(96, 224), (32, 384), (128, 384)]:
        u8 = torch.randint(0, 256, (B, R, R, 3), dtype=torch.uint8)
        d_u8 = u8.cuda()
        out = torch.empty((B, 3, R, R), dtype=torch.float32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def kernel():
            flush.zero_()          # 256 MB > L2: the input is cold
            lib.image_u8_to_f32(d_u8, feed.CLIP_MEAN, feed.CLIP_STD, out=out)

        ms_k = timed(kernel) - timed(lambda: flush.zero_())
        alg = B * R * R * 15
        p_u8 = u8.pin_memory()
        f32 = torch.empty((B, 3, R, R), dtype=torch.float32).pin_memory()
        ms_u8 = timed(lambda: lib.image_u8_to_f32(p_u8.to("cuda", non_blocking=True), feed.CLIP_MEAN, feed.CLIP_STD, out=out))
        ms_f32 = timed(lambda: out.copy_(f32, non_blocking=True))
        rows.append(dict(B=B, res=R, kernel_us=ms_k * 1e3, algorithmic_bytes=alg, achieved_gbps=alg / ms_k * 1e-6, hbm_peak_gbps=hbm,
                         frac=alg / ms_k * 1e-6 / hbm, u8_h2d_plus_kernel_ms=ms_u8, f32_h2d_ms=ms_f32, h2d_bytes_u8=u8.numel(),
                         h2d_bytes_f32=f32.numel() * 4))
    rows += resize_rows()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "feed.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
            print(json.dumps(r))


if __name__ == "__main__":
    main()
