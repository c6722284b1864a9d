"""Correctness + timing of the CTA-pair (cta_group::2) GEMM against the single-CTA kernel and torch."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib as L  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*s, dt=torch.bfloat16):
    return (torch.randn(*s, device="cuda", generator=g) * 0.5).to(dt)


def check(name, M, N, K, a_t=False, b_t=False, **kw):
    A, B = rnd(M, K), rnd(N, K)
    a = A.t().contiguous() if a_t else A
    b = B.t().contiguous() if b_t else B
    ref = A.float() @ B.float().t()
    out_dtype = kw.pop("out_dtype", torch.bfloat16)
    if kw.get("accumulate"):
        o2 = torch.zeros(M, N, device="cuda")
        L.gemm(a, b, a_t=a_t, b_t=b_t, out=o2, block_n=512, **kw)
    else:
        o2 = L.gemm(a, b, a_t=a_t, b_t=b_t, block_n=512, out_dtype=out_dtype, **kw)
    torch.cuda.synchronize()
    err = float((o2.float() - ref).abs().max()) / float(ref.abs().max())
    print(json.dumps(dict(case=name, M=M, N=N, K=K, a_t=a_t, b_t=b_t, rel_err=round(err, 5), ok=err < 1e-2)), flush=True)


def bench(name, M, N, K, a_t=False, b_t=False, reps=20, **kw):
    A, B = rnd(M, K), rnd(N, K)
    a = A.t().contiguous() if a_t else A
    b = B.t().contiguous() if b_t else B
    res = {}
    for tag, bn in (("single", 0), ("pair", 512)):
        fn = lambda: L.gemm(a, b, a_t=a_t, b_t=b_t, block_n=bn, **kw)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        res[tag + "_us"] = round(us, 1)
        res[tag + "_tflops"] = round(2.0 * M * N * K / us / 1e6, 1)
    print(json.dumps(dict(bench=name, M=M, N=N, K=K, **res)), flush=True)


check("small", 256, 256, 128)
check("one_pair_longk", 256, 512, 3072)
check("odd_mtiles", 384, 256, 256)
check("tails", 788, 200, 72)
check("many", 4000, 768, 768)
check("b_mn", 512, 512, 256, b_t=True)
check("ab_mn", 512, 768, 1024, a_t=True, b_t=True)
check("wgrad_splitk", 768, 768, 4000, a_t=True, b_t=True, accumulate=True, split_k=4, out_dtype=torch.float32)
check("vit_fc1", 18912, 3072, 768)
bench("vit_qkv", 18912, 2304, 768)
bench("vit_fc1", 18912, 3072, 768)
bench("vit_fc2", 18912, 768, 3072)
bench("vit_dgrad", 18912, 768, 3072, b_t=True)
bench("txt_ffn1", 3840, 3072, 768)
bench("square_8k", 8192, 8192, 8192, reps=5)
