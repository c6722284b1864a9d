"""Developer check of the tcgen05 GEMM on a B200: correctness matrix + timing vs cuBLAS.
Writes one JSON line per case to gpurun_out/dev_gemm.jsonl (flushed per case so a hang still leaves evidence)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
OUT = open("gpurun_out/dev_gemm.jsonl", "w")


def emit(**kw):
    OUT.write(json.dumps(kw) + "\n")
    OUT.flush()
    print(kw, flush=True)


def gelu(x):
    return torch.nn.functional.gelu(x)


def gelu_grad(x):
    cdf = 0.5 * (1 + torch.erf(x * 0.7071067811865476))
    pdf = torch.exp(-0.5 * x * x) * 0.3989422804014327
    return cdf + x * pdf


def case(name, M, N, K, a_t=False, b_t=False, bn=0, bias=False, act=0, aux_out=False, col_scale=False, rgs=0,
         residual=None, out_dtype=torch.bfloat16, accumulate=False, split_k=1):
    g = torch.Generator(device="cuda").manual_seed(hash(name) % (1 << 31))
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    a_arg = A.t().contiguous() if a_t else A
    b_arg = B.t().contiguous() if b_t else B
    ref = A.float() @ B.float().t()
    kw = {}
    if bias:
        bv = torch.randn(N, device="cuda", generator=g)
        kw["bias"] = bv
        ref = ref + bv
    pre = ref.clone()
    if aux_out:
        kw["aux_out"] = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    if act == 1:
        ref = gelu(ref)
    elif act == 2:
        h = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
        kw["aux_in"] = h
        ref = ref * gelu_grad(h.float())
    elif act == 3:
        ref = torch.tanh(ref)
    if col_scale:
        cs = torch.randn(N, device="cuda", generator=g)
        kw["col_scale"] = cs
        ref = ref * cs
    if rgs:
        ng = (M + rgs - 1) // rgs
        rs = torch.rand(ng, device="cuda", generator=g)
        kw["row_group_scale"] = rs
        kw["rows_per_group"] = rgs
        ref = ref * rs.repeat_interleave(rgs)[:M, None]
    if residual is not None:
        r = torch.randn(M, N, device="cuda", generator=g).to(residual)
        kw["residual"] = r
        ref = ref + r.float()
    out = None
    if accumulate:
        out = torch.randn(M, N, device="cuda", generator=g)
        ref = ref + out
        kw["out"] = out
    try:
        res = lib.gemm(a_arg, b_arg, a_t=a_t, b_t=b_t, act=act, out_dtype=out_dtype, accumulate=accumulate,
                       split_k=split_k, block_n=bn, **kw)
        torch.cuda.synchronize()
        err = (res.float() - ref).abs().max().item()
        scale = ref.abs().max().item()
        aux_err = None
        if aux_out:
            aux_err = (kw["aux_out"].float() - pre).abs().max().item()
        ok = err <= 2e-2 * max(scale, 1.0) and (aux_err is None or aux_err <= 2e-2 * max(pre.abs().max().item(), 1.0))
        emit(case=name, M=M, N=N, K=K, a_t=a_t, b_t=b_t, bn=bn, ok=bool(ok), err=err, scale=scale, aux_err=aux_err)
        return ok
    except Exception as e:  # noqa: BLE001
        emit(case=name, ok=False, error=repr(e))
        return False


def bench(name, M, N, K, a_t=False, b_t=False, bn=0, split_k=1, out_dtype=torch.bfloat16, iters=20, **kw):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    a_arg = A.t().contiguous() if a_t else A
    b_arg = B.t().contiguous() if b_t else B
    out = torch.zeros(M, N, device="cuda", dtype=out_dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    acc = split_k > 1

    def run_mine():
        lib.gemm(a_arg, b_arg, a_t=a_t, b_t=b_t, out=out, accumulate=acc, split_k=split_k, block_n=bn, **kw)

    def run_cublas():
        torch.matmul(A, B.t())

    res = {}
    for label, fn in (("mine", run_mine), ("cublas", run_cublas)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        res[label + "_us"] = med * 1e3
        res[label + "_tflops"] = 2.0 * M * N * K / (med * 1e-3) / 1e12
    emit(bench=name, M=M, N=N, K=K, a_t=a_t, b_t=b_t, bn=bn, split_k=split_k, **res)


def main():
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), flush=True)
    allok = True
    # --- descriptor / pipeline basics
    for bn in (64, 128, 256):
        allok &= case(f"tn_basic_bn{bn}", 256, 256, 128, bn=bn)
    allok &= case("tn_long_k", 256, 512, 3072, bn=256)
    allok &= case("tn_tails", 788, 200, 72, bn=128)
    allok &= case("tn_tiny_n", 96, 2, 1536)
    allok &= case("tn_many_tiles", 4000, 768, 768)
    # --- MN-major operands
    for bn in (64, 128, 256):
        allok &= case(f"b_mn_bn{bn}", 256, 256, 128, b_t=True, bn=bn)
        allok &= case(f"a_mn_bn{bn}", 256, 256, 128, a_t=True, bn=bn)
        allok &= case(f"ab_mn_bn{bn}", 256, 256, 128, a_t=True, b_t=True, bn=bn)
    allok &= case("b_mn_tails", 788, 200, 72, b_t=True)
    allok &= case("ab_mn_tails", 200, 328, 788, a_t=True, b_t=True)
    # --- epilogues
    allok &= case("bias", 300, 768, 768, bias=True)
    allok &= case("gelu_aux", 300, 3072, 768, bias=True, act=1, aux_out=True)
    allok &= case("dgelu", 300, 3072, 768, act=2, b_t=True)
    allok &= case("tanh", 300, 768, 768, bias=True, act=3)
    allok &= case("layerscale_res_f32", 394, 768, 768, bias=True, col_scale=True, rgs=197, aux_out=True,
                  residual=torch.float32, out_dtype=torch.float32)
    allok &= case("res_bf16", 300, 768, 768, bias=True, residual=torch.bfloat16)
    allok &= case("f32_out_tail", 150, 1000, 768, bias=True, out_dtype=torch.float32)
    allok &= case("wgrad_splitk", 768, 768, 4000, a_t=True, b_t=True, out_dtype=torch.float32, accumulate=True,
                  split_k=8)
    allok &= case("wgrad_acc_nosplit", 768, 3072, 520, a_t=True, b_t=True, out_dtype=torch.float32, accumulate=True)
    emit(summary="correctness", all_ok=bool(allok))
    # --- timing (ViT-B shapes at B=96: M = 96*197)
    M = 96 * 197
    bench("vit_qkv", M, 2304, 768)
    bench("vit_proj", M, 768, 768)
    bench("vit_fc1", M, 3072, 768)
    bench("vit_fc2", M, 768, 3072)
    bench("vit_fc1_bn128", M, 3072, 768, bn=128)
    bench("vit_dgrad_fc1", M, 768, 3072, b_t=True)
    bench("vit_wgrad_fc1", 3072, 768, M, a_t=True, b_t=True, split_k=2, out_dtype=torch.float32)
    bench("vit_wgrad_proj", 768, 768, M, a_t=True, b_t=True, split_k=8, out_dtype=torch.float32)
    bench("txt_qkv", 96 * 40, 2304, 768)
    bench("txt_ffn1", 96 * 40, 3072, 768)
    bench("mlm_dec", 1440, 50265, 768, out_dtype=torch.float32)
    bench("square_8k", 8192, 8192, 8192)
    emit(summary="done", launches=lib.launch_count())


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("elapsed", time.time() - t0)
