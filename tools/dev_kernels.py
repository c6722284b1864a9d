"""Developer timing of the non-GEMM kernels at the pre-training shapes (B=96/GPU).
One JSON line per kernel to gpurun_out/dev_kernels.jsonl.  Timed with CUDA events over `reps` launches that rotate over
`nbuf` independent buffer sets (so consecutive launches do not hit the same L2-resident data).

    python tools/dev_kernels.py [attn|ln|all]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib as L  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
OUT = open("gpurun_out/dev_kernels.jsonl", "a")
HBM = 6539.9


def emit(**kw):
    OUT.write(json.dumps(kw) + "\n")
    OUT.flush()
    print(kw, flush=True)


def timeit(fns, reps=20):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def attn_case(name, B, H, Lq, Lk, mode, Bkv=None, p=0.0, nbuf=3):
    D = H * 64
    dev = "cuda"
    Bkv = B if Bkv is None else Bkv
    sets = []
    for i in range(nbuf):
        g = torch.Generator(device=dev).manual_seed(i)
        if mode == "cross":
            q = torch.randn(B * Lq, D, device=dev, generator=g).bfloat16()
            kv = torch.randn(Bkv * Lk, 2 * D, device=dev, generator=g).bfloat16()
            k, v = kv[:, :D], kv[:, D:]
            dq, dkv = torch.empty_like(q), torch.empty_like(kv)
            dk, dv = dkv[:, :D], dkv[:, D:]
            kv_index = (torch.arange(B, device=dev) % Bkv).to(torch.int32)
            order = torch.argsort(kv_index.long(), stable=True).to(torch.int32)
            offs = torch.zeros(Bkv + 1, dtype=torch.int32, device=dev)
            offs[1:] = torch.cumsum(torch.bincount(kv_index.long(), minlength=Bkv), 0).to(torch.int32)
        else:
            qkv = torch.randn(B * Lq, 3 * D, device=dev, generator=g).bfloat16()
            q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
            dqkv = torch.empty_like(qkv)
            dq, dk, dv = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
            kv_index = order = offs = None
        ld = (Lk + 7) // 8 * 8
        bias = torch.randn(H, Lq, ld, device=dev, generator=g) if mode in ("vit", "vit_tc") else None
        table = torch.randn(732, H, device=dev, generator=g) if mode == "vit_tc" else None
        kmask = torch.zeros(B, Lk, device=dev) if mode in ("text", "text_tc") else None
        dout = torch.randn(B * Lq, D, device=dev, generator=g).bfloat16()
        ds = torch.empty(B, H, Lq, ld, device=dev, dtype=torch.bfloat16) if mode in ("vit", "vit_tc") else None
        out = torch.empty(B * Lq, D, device=dev, dtype=torch.bfloat16)
        sets.append(dict(q=q, k=k, v=v, dq=dq, dk=dk, dv=dv, kv_index=kv_index, order=order, offs=offs, bias=bias, kmask=kmask,
                         table=table, dtab=None if table is None else torch.zeros_like(table),
                         dout=dout, ds=ds, out=out))
    lses = []

    def fwd(s):
        o, lse = L.attention_fwd(s["q"], s["k"], s["v"], B, H, Lq, Lk, 0.125, Bkv=Bkv, bias=s["bias"], kmask=s["kmask"],
                                 kv_index=s["kv_index"], dropout_p=p, dropout_seed=7, out=s["out"], rel_table=s["table"],
                                 rel_window=14 if s["table"] is not None else 0, allow_tc=mode in ("vit_tc", "plain_tc", "cross", "text_tc"),
                                 kv_offsets=s["offs"], kv_samples=s["order"])
        s["lse"] = lse

    tc = mode in ("vit_tc", "plain_tc", "cross", "text_tc")

    def bwd(s):
        L.attention_bwd(s["dout"], s["q"], s["k"], s["v"], s["out"], s["lse"], B, H, Lq, Lk, 0.125, s["dq"], s["dk"], s["dv"],
                        Bkv=Bkv, bias=s["bias"], kmask=s["kmask"], kv_index=s["kv_index"], kv_offsets=s["offs"],
                        kv_samples=s["order"], dropout_p=p, dropout_seed=7, ds_dump=s["ds"],
                        rel_table=s["table"], rel_window=14 if s["table"] is not None else 0,
                        allow_tc=tc)

    us_f = timeit([lambda s=s: fwd(s) for s in sets])
    us_b = timeit([lambda s=s: bwd(s) for s in sets])
    flops_f = 4.0 * B * H * Lq * Lk * 64
    emit(kernel="attention", case=name, B=B, H=H, Lq=Lq, Lk=Lk, Bkv=Bkv, dropout=p, fwd_us=round(us_f, 1), bwd_us=round(us_b, 1),
         fwd_tflops=round(flops_f / us_f / 1e6, 1), bwd_tflops=round(2.5 * flops_f / us_b / 1e6, 1))


def ln_case(M, D=768, nbuf=3):
    dev = "cuda"
    xs = [torch.randn(M, D, device=dev) for _ in range(nbuf)]
    dys = [torch.randn(M, D, device=dev).bfloat16() for _ in range(nbuf)]
    adds = [torch.randn(M, D, device=dev) for _ in range(nbuf)]
    w, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    dw, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    stats = []
    for x in xs:
        _, st, _ = L.layernorm_fwd(x, w, b, 1e-6)
        stats.append(st)
    us_f = timeit([lambda x=x: L.layernorm_fwd(x, w, b, 1e-6) for x in xs])
    us_f2 = timeit([lambda x=x: L.layernorm_fwd(x, w, b, 1e-5, want_f32_copy=True) for x in xs])
    us_b = timeit([lambda i=i: L.layernorm_bwd(dys[i], xs[i], stats[i], w, dw, db, add_in=adds[i]) for i in range(nbuf)])
    bf, bb = M * D * (4 + 2), M * D * (2 + 4 + 4 + 4)
    emit(kernel="layernorm", M=M, D=D, fwd_us=round(us_f, 1), fwd_gbs=round(bf / us_f / 1e3), fwd_f32copy_us=round(us_f2, 1),
         fwd_f32copy_gbs=round((bf + M * D * 4) / us_f2 / 1e3), bwd_us=round(us_b, 1), bwd_gbs=round(bb / us_b / 1e3), peak_gbs=HBM)


def misc_case(M=18912, D=768):
    dev = "cuda"
    x16 = [torch.randn(M, 3 * D, device=dev).bfloat16() for _ in range(3)]
    out = torch.zeros(3 * D, device=dev)
    us = timeit([lambda x=x: L.colsum_into(x, out) for x in x16])
    emit(kernel="colsum", M=M, N=3 * D, us=round(us, 1), gbs=round(M * 3 * D * 2 / us / 1e3))
    dx = [torch.randn(M, D, device=dev) for _ in range(3)]
    z = [torch.randn(M, D, device=dev).bfloat16() for _ in range(3)]
    gam, dg, dbias = torch.ones(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    us = timeit([lambda i=i: L.layerscale_bwd(dx[i], z[i], gam, dg, dbias) for i in range(3)])
    emit(kernel="layerscale_bwd", M=M, D=D, us=round(us, 1), gbs=round(M * D * (4 + 2 + 2) / us / 1e3))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.cuda.set_device(0)
    if what == "tc":
        attn_case("vit_self_tcgen05", 96, 12, 197, 197, "vit_tc")
        attn_case("fusion_cross", 384, 12, 40, 197, "cross", Bkv=96, p=0.1)
    if what == "self":
        attn_case("text_self", 96, 12, 40, 40, "text", p=0.1)
        attn_case("fusion_self", 384, 12, 40, 40, "text", p=0.1)
        attn_case("text_self_tcgen05", 96, 12, 40, 40, "text_tc", p=0.1)
        attn_case("fusion_self_tcgen05", 384, 12, 40, 40, "text_tc", p=0.1)
    if what in ("attn", "all"):
        attn_case("vit_self_tcgen05", 96, 12, 197, 197, "vit_tc")
        attn_case("vqkd_self_tcgen05", 96, 12, 197, 197, "plain_tc")
        attn_case("vit_self", 96, 12, 197, 197, "vit")
        attn_case("vqkd_self", 96, 12, 197, 197, "plain")
        attn_case("text_self", 96, 12, 40, 40, "text", p=0.1)
        attn_case("fusion_self", 384, 12, 40, 40, "text", p=0.1)
        attn_case("text_self_tcgen05", 96, 12, 40, 40, "text_tc", p=0.1)
        attn_case("fusion_self_tcgen05", 384, 12, 40, 40, "text_tc", p=0.1)
        attn_case("fusion_cross", 384, 12, 40, 197, "cross", Bkv=96, p=0.1)
    if what in ("ln", "all"):
        ln_case(18912)
        ln_case(3840)
        ln_case(15360)
        misc_case()
