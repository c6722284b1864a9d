"""Tile-shape sweep of the tcgen05 GEMM on the shapes whose tile count is below / near one wave (text encoder, fusion
text side): which of the CTA-pair 256x256 kernel (block_n=512) and the single-CTA 128 x {256,128,64} kernels is fastest.
    python tools/gemm_sweep.py  -> gpurun_out/gemm_sweep.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib as L  # noqa: E402


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    big = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(n):
        big.zero_()   # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def main():
    L.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    for M in (1440, 3840, 7680, 15360):
        for N, K in ((768, 768), (2304, 768), (3072, 768), (768, 3072)):
            a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
            w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
            bias = torch.randn(N, device="cuda", generator=g)
            res = torch.randn(M, N, device="cuda", generator=g)
            for mode in ("plain_bf16", "res_f32"):
                rec = dict(M=M, N=N, K=K, mode=mode)
                for bn in (512, 256, 128, 64):
                    if bn == 512 and M < 256:
                        continue
                    try:
                        if mode == "plain_bf16":
                            fn = lambda: L.gemm(a, w, bias=bias, block_n=bn)
                        else:
                            fn = lambda: L.gemm(a, w, bias=bias, residual=res, out_dtype=torch.float32, block_n=bn)
                        us = timeit(fn)
                        rec[f"bn{bn}_us"] = round(us, 1)
                        rec[f"bn{bn}_tflops"] = round(2.0 * M * N * K / us / 1e6, 0)
                    except Exception as e:
                        rec[f"bn{bn}_err"] = str(e)[:80]
                rows.append(rec)
                print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_sweep.jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
