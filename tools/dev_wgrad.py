"""Sweep kernel / split-K choices for the weight-gradient GEMMs dW[Nout, Kin] += dY[T, Nout]^T X[T, Kin] of the step.
    python tools/dev_wgrad.py            (prints one JSON line per shape: best and all timings)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib as L  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
shapes = [(768, 768, 18912), (768, 768, 15360), (768, 768, 3840), (768, 3072, 18912), (3072, 768, 18912), (2304, 768, 18912),
          (2304, 768, 3840), (1536, 768, 18912), (768, 3072, 3840), (3072, 768, 3840)]
for n_out, k_in, T in shapes:
    dy = (torch.randn(T, n_out, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    x = (torch.randn(T, k_in, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    out = torch.zeros(n_out, k_in, device=dev)
    res = {}
    for bn in (0, 128, 256, 512):
        for sk in (1, 2, 3, 4, 6, 8, 12, 16, 24):
            fn = lambda: L.gemm(dy, x, a_t=True, b_t=True, out=out, accumulate=True, split_k=sk, block_n=bn)
            try:
                for _ in range(2):
                    fn()
            except RuntimeError:
                continue
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[f"bn{bn}_sk{sk}"] = round(e0.elapsed_time(e1) * 100, 1)
    best = min(res, key=res.get)
    top = dict(sorted(res.items(), key=lambda kv: kv[1])[:6])
    print(json.dumps(dict(shape=[n_out, k_in, T], best=best, us=res[best], tflops=round(2.0 * n_out * k_in * T / res[best] / 1e6), top=top)),
          flush=True)
