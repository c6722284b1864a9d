"""ctypes binding of libxfm_b200.so (the C-ABI declared in include/xfm_b200.h).

The library is the only compute path of this package: there is no CPU or eager-PyTorch
fallback.  If the shared object is missing, or a GPU is required and absent, calls raise.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxfm_b200.so")

_lib = None
_inited_devices = set()


class GemmParams(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("c_dtype", C.c_int32), ("split_k", C.c_int32), ("accumulate", C.c_int32),
        ("act", C.c_int32), ("res_dtype", C.c_int32), ("rows_per_group", C.c_int32),
        ("block_n", C.c_int32),
        ("bias", C.c_void_p), ("aux_in", C.c_void_p), ("aux_out", C.c_void_p),
        ("ld_aux_in", C.c_int64), ("ld_aux_out", C.c_int64), ("ld_res", C.c_int64),
        ("col_scale", C.c_void_p), ("row_group_scale", C.c_void_p), ("residual", C.c_void_p),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64),
    ]


def load():
    """Load the shared library (no CUDA context needed).  Raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(xfm_b200 has no fallback path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.xfm_version.restype = C.c_int
        _lib.xfm_init.restype = C.c_int
        _lib.xfm_launch_count.restype = C.c_int64
        _lib.xfm_last_error.restype = C.c_char_p
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().xfm_last_error().decode()
        raise RuntimeError(f"xfm_b200: {what} failed with code {rc}: {msg}")


def lib():
    """Library handle, initialised for the current CUDA device."""
    l = load()
    if not torch.cuda.is_available():
        raise RuntimeError("xfm_b200 needs a CUDA device (sm_100a); there is no CPU path")
    dev = torch.cuda.current_device()
    if dev not in _inited_devices:
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make sure the primary context exists
        check(l.xfm_init(), "xfm_init")
        _inited_devices.add(dev)
    return l


def launch_count():
    return int(load().xfm_launch_count())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_DT = {torch.bfloat16: 0, torch.float32: 1}


def gemm(a, b, *, a_t=False, b_t=False, out=None, out_dtype=torch.bfloat16, bias=None, act=0, aux_in=None,
         aux_out=None, col_scale=None, row_group_scale=None, rows_per_group=1, residual=None, dropout_p=0.0,
         dropout_seed=0, accumulate=False, split_k=1, block_n=0):
    """C[M,N] = epilogue(A · B^T).

    a: bf16 [M,K] (or [K,M] with a_t=True: the stored tensor is the transpose, "MN-major").
    b: bf16 [N,K] (or [K,N] with b_t=True).
    """
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_t:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_t:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, a_t, b_t)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    p = GemmParams()
    p.A, p.B, p.C = a.data_ptr(), b.data_ptr(), out.data_ptr()
    p.lda, p.ldb, p.ldc = a.stride(0), b.stride(0), out.stride(0)
    p.M, p.N, p.K = M, N, K
    p.a_mn_major, p.b_mn_major = int(a_t), int(b_t)
    p.c_dtype = _DT[out.dtype]
    p.split_k = split_k
    p.accumulate = int(accumulate)
    p.act = act
    p.block_n = block_n
    p.rows_per_group = rows_per_group
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        p.bias = bias.data_ptr()
    if aux_in is not None:
        assert aux_in.dtype == torch.bfloat16 and aux_in.shape == (M, N)
        p.aux_in, p.ld_aux_in = aux_in.data_ptr(), aux_in.stride(0)
    if aux_out is not None:
        assert aux_out.dtype == torch.bfloat16 and aux_out.shape == (M, N)
        p.aux_out, p.ld_aux_out = aux_out.data_ptr(), aux_out.stride(0)
    if col_scale is not None:
        assert col_scale.dtype == torch.float32 and col_scale.numel() == N
        p.col_scale = col_scale.data_ptr()
    if row_group_scale is not None:
        assert row_group_scale.dtype == torch.float32
        p.row_group_scale = row_group_scale.data_ptr()
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1
        p.residual, p.ld_res, p.res_dtype = residual.data_ptr(), residual.stride(0), _DT[residual.dtype]
    p.dropout_p = dropout_p
    p.dropout_seed = dropout_seed
    check(lib().xfm_gemm_bf16(C.byref(p), stream_ptr()), "xfm_gemm_bf16")
    return out
