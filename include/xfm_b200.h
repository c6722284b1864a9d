/*
 * xfm_b200 — C-ABI of the B200-native XFM hot path (libxfm_b200.so).
 *
 * Every entry point takes plain device pointers, sizes and a cudaStream_t (passed as void*);
 * no torch types cross this boundary.  All functions return 0 on success or a non-zero
 * cudaError_t / negative xfm error code; the Python host side (xfm_b200/lib.py) raises on != 0.
 *
 * The reference (zhangxinsong-nlp/XFM) is pure PyTorch: the "FFI" this library replaces is the
 * set of ATen calls made from the modules below (file:line relative to the reference root).
 * Each declaration cites the reference code whose arithmetic it reproduces.
 *
 * Conventions: activations are row-major [rows, features]; "bf16" = __nv_bfloat16; "f32" = float.
 */
#ifndef XFM_B200_H
#define XFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XFM_ERR_BAD_ARG (-2)
#define XFM_ERR_NO_DRIVER (-3)

/* ------------------------------------------------------------------------------------------------
 * Library
 * ---------------------------------------------------------------------------------------------- */
/* Returns the ABI version (monotonic integer). */
int xfm_version(void);
/* Resolves driver entry points and sets kernel attributes on the current device.  Must be called
 * once per process after the CUDA context exists. */
int xfm_init(void);
/* Number of kernels launched by this library since process start (all entry points). */
int64_t xfm_launch_count(void);
/* Human-readable description of the last error recorded by this library on this thread. */
const char* xfm_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * K1 — tcgen05 GEMM with fused epilogue.
 *   C[M,N] = epilogue( A[M,K] · B[N,K]^T )
 * Replaces torch addmm / F.linear at: beit2.py:132,160 (qkv, proj), beit2.py:62-69 (Mlp),
 * xroberta.py:211,224-225,233-234 (query/key/value), :301 (SelfOutput.dense), :368 (Intermediate),
 * :382 (Output.dense), :1325-1333 (LM head), xfm.py:115-121,500-501 (heads), model_vqkd.py:86-90,
 * and their autograd backward (dgrad / wgrad) GEMMs.
 *
 * Operand layouts (bf16):
 *   a_mn_major == 0 : A stored [M, K] row-major, leading dimension lda (elements).
 *   a_mn_major == 1 : A stored [K, M] row-major (i.e. the transpose is in memory), lda = row stride.
 *   b_mn_major == 0 : B stored [N, K] row-major (a torch Linear weight), ldb.
 *   b_mn_major == 1 : B stored [K, N] row-major, ldb.
 * Leading dimensions must be multiples of 8 elements and base pointers 16-byte aligned (TMA).
 *
 * Epilogue, applied per element in this order (each step optional):
 *   v  = acc (+ bias[n])
 *   aux_out[m,n] = bf16(v)                         (pre-activation / pre-scale copy for backward)
 *   v  = act(v)            act: 0 none | 1 GELU(erf) | 2 v * GELU'(aux_in[m,n]) | 3 tanh
 *   v *= col_scale[n]                              (BEiT LayerScale gamma, beit2.py:204-205)
 *   v *= row_group_scale[m / rows_per_group]       (DropPath keep/keep_prob per sample)
 *   v  = dropout(v; p, seed, element index)        (xroberta.py:302,383)
 *   v += residual[m,n]                             (bf16 or f32)
 *   C[m,n] (+)= v                                  (bf16 or f32; accumulate / split-K use f32 atomics)
 * ---------------------------------------------------------------------------------------------- */
typedef struct xfm_gemm_params {
  const void* A;
  const void* B;
  void* C;
  int64_t lda, ldb, ldc;
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  int32_t c_dtype;    /* 0 = bf16, 1 = f32 */
  int32_t split_k;    /* >= 1; > 1 requires c_dtype == 1 and accumulate == 1 (atomic f32 adds) */
  int32_t accumulate; /* 0: C = v, 1: C += v (f32 only) */
  int32_t act;        /* 0 none, 1 gelu, 2 dgelu(aux_in), 3 tanh */
  int32_t res_dtype;  /* 0 = bf16, 1 = f32 */
  int32_t rows_per_group;
  int32_t block_n;    /* 0 = auto, else 64 / 128 / 256 */
  const float* bias;
  const void* aux_in;
  void* aux_out;
  int64_t ld_aux_in, ld_aux_out, ld_res;
  const float* col_scale;
  const float* row_group_scale;
  const void* residual;
  float dropout_p;
  uint64_t dropout_seed;
} xfm_gemm_params;

int xfm_gemm_bf16(const xfm_gemm_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XFM_B200_H */
