"""Time (and give ncu something short to profile) the fused-epilogue GEMM variants at the ViT-B shapes of the step.
    python tools/gemm_case.py [case ...]      cases: plain gelu dgelu proj fc2 wgrad ffn_out all
                                              (+ gelu_noaux plain_aux plain_wide: the fc1 epilogue taken apart)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib as L  # noqa: E402

dev = "cuda"
M, D, F = 18912, 768, 3072
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, dtype=torch.bfloat16):
    return (torch.randn(*shape, device=dev, generator=g) * 0.5).to(dtype)


def build(case):
    if case == "plain":
        a, w, b = rnd(M, D), rnd(3 * D, D), rnd(3 * D, dtype=torch.float32)
        return (lambda: L.gemm(a, w, bias=b)), 2.0 * M * 3 * D * D, M * 3 * D * 2
    if case == "gelu":
        a, w, b, aux = rnd(M, D), rnd(F, D), rnd(F, dtype=torch.float32), torch.empty(M, F, device=dev, dtype=torch.bfloat16)
        return (lambda: L.gemm(a, w, bias=b, act=1, aux_out=aux)), 2.0 * M * F * D, M * F * 4
    if case == "gelu_noaux":   # the GELU arithmetic without the second (saved pre-activation) store
        a, w, b = rnd(M, D), rnd(F, D), rnd(F, dtype=torch.float32)
        return (lambda: L.gemm(a, w, bias=b, act=1)), 2.0 * M * F * D, M * F * 2
    if case == "plain_aux":    # the second store without the GELU arithmetic
        a, w, b, aux = rnd(M, D), rnd(F, D), rnd(F, dtype=torch.float32), torch.empty(M, F, device=dev, dtype=torch.bfloat16)
        return (lambda: L.gemm(a, w, bias=b, aux_out=aux)), 2.0 * M * F * D, M * F * 4
    if case == "plain_wide":   # bias only, N = 3072 (same tile count as the GELU case)
        a, w, b = rnd(M, D), rnd(F, D), rnd(F, dtype=torch.float32)
        return (lambda: L.gemm(a, w, bias=b)), 2.0 * M * F * D, M * F * 2
    if case == "dgelu":
        a, w, aux = rnd(M, D), rnd(D, F), rnd(M, F)
        return (lambda: L.gemm(a, w, b_t=True, act=2, aux_in=aux)), 2.0 * M * F * D, M * F * 4
    if case in ("proj", "fc2"):
        K = D if case == "proj" else F
        a, w, b = rnd(M, K), rnd(D, K), rnd(D, dtype=torch.float32)
        res, aux, cs = rnd(M, D, dtype=torch.float32), torch.empty(M, D, device=dev, dtype=torch.bfloat16), rnd(D, dtype=torch.float32)
        rgs = torch.ones(96, device=dev)
        return (lambda: L.gemm(a, w, bias=b, col_scale=cs, row_group_scale=rgs, rows_per_group=197, residual=res, aux_out=aux,
                               out_dtype=torch.float32)), 2.0 * M * D * K, M * D * 10
    if case == "ffn_out":   # roberta output dense: dropout + f32 residual -> f32
        Mt = 15360
        a, w, b, res = rnd(Mt, F), rnd(D, F), rnd(D, dtype=torch.float32), rnd(Mt, D, dtype=torch.float32)
        return (lambda: L.gemm(a, w, bias=b, dropout_p=0.1, dropout_seed=5, residual=res, out_dtype=torch.float32)), 2.0 * Mt * D * F, Mt * D * 8
    if case == "wgrad":
        dy, x, out = rnd(M, F), rnd(M, D), torch.zeros(F, D, device=dev)
        return (lambda: L.gemm(dy, x, a_t=True, b_t=True, out=out, accumulate=True, split_k=2)), 2.0 * M * F * D, F * D * 4
    raise SystemExit(f"unknown case {case}")


cases = sys.argv[1:] or ["all"]
reps = int(os.environ.get("REPS", "20"))
if cases and cases[0].startswith("reps="):
    reps = int(cases.pop(0)[5:])
if not cases or cases == ["all"]:
    cases = ["plain", "gelu", "dgelu", "proj", "fc2", "ffn_out", "wgrad"]
for c in cases:
    fn, flops, epi_bytes = build(c)
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
    print(json.dumps(dict(case=c, us=round(us, 1), tflops=round(flops / us / 1e6, 1), epilogue_gbs=round(epi_bytes / us / 1e3))), flush=True)
