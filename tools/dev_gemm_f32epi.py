"""Timing + correctness of the fp32-residual GEMM epilogue (N = K = 768 projections of the step) — run once per setting of
XFM_GEMM_F32_DEEP (0 = two-tile ring / 5 stages, default = four-tile ring / 3 stages for K <= 1024).

    python tools/dev_gemm_f32epi.py [tag]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib as L  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("XFM_GEMM_F32_DEEP", "default")
g = torch.Generator(device="cuda").manual_seed(0)
os.makedirs("gpurun_out", exist_ok=True)
OUT = open("gpurun_out/dev_gemm_f32epi.jsonl", "a")


def rnd(*s, dt=torch.bfloat16, scale=0.5):
    return (torch.randn(*s, device="cuda", generator=g) * scale).to(dt)


def case(name, M, N, K, layerscale=False, drop=0.0, b_t=False, nbuf=3, reps=30, aux=None, scale=None):
    """aux / scale override which halves of the LayerScale epilogue are on (timing only)."""
    if aux is not None or scale is not None:
        return parts(name, M, N, K, bool(aux), bool(scale), nbuf, reps)
    sets = []
    for i in range(nbuf):
        A, B = rnd(M, K), rnd(N, K, scale=0.05)
        b = B.t().contiguous() if b_t else B
        res = rnd(M, N, dt=torch.float32)
        sets.append((A, B, b, res))
    bias = rnd(N, dt=torch.float32)
    gamma = rnd(N, dt=torch.float32) if layerscale else None
    rs = (torch.rand(M // 197 + 1, device="cuda", generator=g) + 0.5) if layerscale else None
    kw = dict(bias=bias, out_dtype=torch.float32, dropout_p=drop, dropout_seed=1234)
    if layerscale:
        kw.update(col_scale=gamma, row_group_scale=rs, rows_per_group=197)

    def run(i):
        A, B, b, res = sets[i % nbuf]
        aux = torch.empty(M, N, dtype=torch.bfloat16, device="cuda") if layerscale else None
        return L.gemm(A, b, b_t=b_t, residual=res, aux_out=aux, **kw), aux
    out, aux = run(0)
    torch.cuda.synchronize()
    A, B, b, res = sets[0]
    z = A.float() @ B.float().t() + bias
    err = None
    if drop == 0.0:
        ref = z * (gamma * rs.repeat_interleave(197)[:M, None] if layerscale else 1.0) + res
        err = float((out - ref).abs().max() / ref.abs().max())
        if layerscale:
            err = max(err, float((aux.float() - z).abs().max() / z.abs().max()) / 8)   # bf16 rounding of z: 2^-8 relative
    else:   # kept elements equal the undropped value / (1 - p); dropped ones equal the residual
        ref_keep = z / (1.0 - drop) + res
        is_keep = (out - ref_keep).abs() <= 1e-2 * ref_keep.abs().max()
        is_drop = (out - res).abs() <= 1e-6
        err = 1.0 - float((is_keep | is_drop).float().mean())
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    byts = M * K * 2 + N * K * 2 + M * N * (4 + 4 + (2 if layerscale else 0))
    rec = dict(tag=tag, case=name, M=M, N=N, K=K, us=round(us, 1), tflops=round(2.0 * M * N * K / us / 1e6, 1),
               gbs=round(byts / us / 1e3), err=err)
    OUT.write(json.dumps(rec) + "\n")
    OUT.flush()
    print(rec, flush=True)


def parts(name, M, N, K, aux, scale, nbuf, reps):
    sets = [(rnd(M, K), rnd(N, K, scale=0.05), rnd(M, N, dt=torch.float32)) for _ in range(nbuf)]
    bias, gamma = rnd(N, dt=torch.float32), rnd(N, dt=torch.float32)
    rs = torch.rand(M // 197 + 1, device="cuda", generator=g) + 0.5
    kw = dict(bias=bias, out_dtype=torch.float32)
    if scale:
        kw.update(col_scale=gamma, row_group_scale=rs, rows_per_group=197)
    auxb = torch.empty(M, N, dtype=torch.bfloat16, device="cuda") if aux else None
    outb = torch.empty(M, N, dtype=torch.float32, device="cuda")

    def run(i):
        A, B, res = sets[i % nbuf]
        L.gemm(A, B, residual=res, aux_out=auxb, out=outb, **kw)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    rec = dict(tag=tag, case=name, M=M, N=N, K=K, aux=aux, scale=scale, us=round(us, 1), tflops=round(2.0 * M * N * K / us / 1e6, 1))
    OUT.write(json.dumps(rec) + "\n")
    OUT.flush()
    print(rec, flush=True)


case("parts_res_only", 18912, 768, 768, aux=False, scale=False)
case("parts_res_aux", 18912, 768, 768, aux=True, scale=False)
case("parts_res_scale", 18912, 768, 768, aux=False, scale=True)
case("parts_res_aux_scale", 18912, 768, 768, aux=True, scale=True)
case("vit_proj_layerscale", 18912, 768, 768, layerscale=True)
case("fusion_out_dense", 15360, 768, 768, drop=0.1)
case("text_out_dense", 7680, 768, 768, drop=0.1)
case("vit_fc2_layerscale_K3072", 18912, 768, 3072, layerscale=True)
case("dgrad_residual", 18912, 768, 768, b_t=True)
