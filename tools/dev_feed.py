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
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "feed.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
            print(json.dumps(r))


if __name__ == "__main__":
    main()
