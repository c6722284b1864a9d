// K1 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue(A[M,K] · B[N,K]^T), bf16 operands, fp32 accumulation in TMEM.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread UMMA issuer,
// warps 2..5 = epilogue (one TMEM lane quadrant each).  Operands are staged by TMA into a
// multi-stage ring of SWIZZLE_128B tiles (128 x 64 for A, BLOCK_N x 64 for B); the accumulator
// is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs
// of tile i+1.  Both operands may be K-major (torch Linear forward) or MN-major (the operand's
// transpose is what lives in memory: dgrad reads W as [K_red, N], wgrad reads dY / X as [tokens, *]).
// Split-K (wgrad: few output tiles, long reduction) accumulates with fp32 vector atomics.
#include "common.cuh"
#include "internal.h"

namespace xfm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

struct GemmArgs {
  int M, N, K;
  int num_m_tiles, num_n_tiles, split_k, kb_total;
  int c_dtype, accumulate, act, res_dtype, rows_per_group;
  int64_t ldc, ld_aux_in, ld_aux_out, ld_res;
  void* C;
  const float* bias;
  const bf16* aux_in;
  bf16* aux_out;
  const float* col_scale;
  const float* row_group_scale;
  const void* residual;
  float dropout_p;
  uint64_t dropout_seed;
  int vec_ok;  // all row pointers 16-byte aligned for 32-column chunks
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

XFM_DEVINL void store_chunk(const GemmArgs& g, int row, int n, int ncols, float (&v)[32]) {
  // v[0..ncols) are final values for C[row, n .. n+ncols)
  if (g.c_dtype == 0) {
    bf16* dst = (bf16*)g.C + (int64_t)row * g.ldc + n;
    if (ncols == 32 && g.vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
        u.x = *(uint32_t*)&t0; u.y = *(uint32_t*)&t1; u.z = *(uint32_t*)&t2; u.w = *(uint32_t*)&t3;
        *(uint4*)(dst + j) = u;
      }
    } else {
      for (int j = 0; j < ncols; ++j) dst[j] = __float2bfloat16(v[j]);
    }
  } else {
    float* dst = (float*)g.C + (int64_t)row * g.ldc + n;
    if (g.accumulate) {
      if (ncols == 32 && g.vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                       "f"(v[j + 2]), "f"(v[j + 3])
                       : "memory");
        }
      } else {
        for (int j = 0; j < ncols; ++j) atomicAdd(dst + j, v[j]);
      }
    } else {
      if (ncols == 32 && g.vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *(float4*)(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        for (int j = 0; j < ncols; ++j) dst[j] = v[j];
      }
    }
  }
}

XFM_DEVINL void load_bf16_chunk(const bf16* src, int ncols, bool vec, float (&o)[32]) {
  if (ncols == 32 && vec) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 u = *(const uint4*)(src + j);
      const __nv_bfloat162* h = (const __nv_bfloat162*)&u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 f = __bfloat1622float2(h[q]);
        o[j + 2 * q] = f.x;
        o[j + 2 * q + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = (j < ncols) ? __bfloat162float(src[j]) : 0.f;
  }
}

XFM_DEVINL void epilogue_chunk(const GemmArgs& g, int row, int n, uint32_t (&r)[32]) {
  const int ncols = min(32, g.N - n);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (g.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += (j < ncols) ? __ldg(g.bias + n + j) : 0.f;
  }
  if (g.aux_out) {
    bf16* dst = g.aux_out + (int64_t)row * g.ld_aux_out + n;
    if (ncols == 32 && g.vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
        u.x = *(uint32_t*)&t0; u.y = *(uint32_t*)&t1; u.z = *(uint32_t*)&t2; u.w = *(uint32_t*)&t3;
        *(uint4*)(dst + j) = u;
      }
    } else {
      for (int j = 0; j < ncols; ++j) dst[j] = __float2bfloat16(v[j]);
    }
  }
  if (g.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (g.act == 2) {
    float a[32];
    load_bf16_chunk(g.aux_in + (int64_t)row * g.ld_aux_in + n, ncols, g.vec_ok, a);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= gelu_erf_grad(a[j]);
  } else if (g.act == 3) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
  }
  if (g.col_scale) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= (j < ncols) ? __ldg(g.col_scale + n + j) : 0.f;
  }
  if (g.row_group_scale) {
    const float s = __ldg(g.row_group_scale + row / g.rows_per_group);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= s;
  }
  if (g.dropout_p > 0.f) {
    const float inv_keep = 1.0f / (1.0f - g.dropout_p);
    const uint64_t base = (uint64_t)row * (uint64_t)g.N + (uint64_t)n;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (hash_uniform(g.dropout_seed, base + j) >= g.dropout_p) ? v[j] * inv_keep : 0.f;
  }
  if (g.residual) {
    if (g.res_dtype == 0) {
      float a[32];
      load_bf16_chunk((const bf16*)g.residual + (int64_t)row * g.ld_res + n, ncols, g.vec_ok, a);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += a[j];
    } else {
      const float* src = (const float*)g.residual + (int64_t)row * g.ld_res + n;
      if (ncols == 32 && g.vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 f = *(const float4*)(src + j);
          v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += (j < ncols) ? src[j] : 0.f;
      }
    }
  }
  store_chunk(g, row, n, ncols, v);
}

template <int BLOCK_N, int A_MN, int B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmArgs g) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = (uint64_t*)(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = g.num_m_tiles * g.num_n_tiles * g.split_k;
  const int kb_per_split = (g.kb_total + g.split_k - 1) / g.split_k;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int ks = t % g.split_k;
        const int rest = t / g.split_k;
        const int m0 = (rest % g.num_m_tiles) * BLOCK_M;
        const int n0 = (rest / g.num_m_tiles) * BLOCK_N;
        const int kb0 = ks * kb_per_split;
        const int kb1 = min(g.kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d(sa + j * (64 * BLOCK_K * 2), &map_a, &full_bar[stage], m0 + 64 * j, kb * BLOCK_K);
          } else {
            tma_load_2d(sa, &map_a, &full_bar[stage], kb * BLOCK_K, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * (64 * BLOCK_K * 2), &map_b, &full_bar[stage], n0 + 64 * j, kb * BLOCK_K);
          } else {
            tma_load_2d(sb, &map_b, &full_bar[stage], kb * BLOCK_K, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int ks = t % g.split_k;
        const int kb0 = ks * kb_per_split;
        const int kb1 = min(g.kb_total, kb0 + kb_per_split);
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: advance 16 elements = 32 B inside the 128 B swizzle row.
            // MN-major: advance 16 k-rows = two 8-row groups = 2048 B.
            const uint64_t a_desc = A_MN ? make_smem_desc(a_addr + k * 2048, 64 * BLOCK_K * 2, 1024)
                                         : make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t b_desc = B_MN ? make_smem_desc(b_addr + k * 2048, 64 * BLOCK_K * 2, 1024)
                                         : make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int rest = t / g.split_k;
      const int m0 = (rest % g.num_m_tiles) * BLOCK_M;
      const int n0 = (rest / g.num_m_tiles) * BLOCK_N;
      const int ks = t % g.split_k;
      const int kb0 = ks * kb_per_split;
      const int kb1 = min(g.kb_total, kb0 + kb_per_split);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N);
      const bool has_work = kb1 > kb0;  // empty split slices contribute nothing
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int n = n0 + c * 32;
        if (n >= g.N) break;
        uint32_t r[32];
        tmem_ld_32x32(t_base + c * 32, r);
        tmem_ld_wait();
        if (row < g.M && has_work) epilogue_chunk(g, row, n, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host
static int encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                     uint32_t box_inner, uint32_t box_outer) {
  auto fn = get_tensor_map_encoder();
  if (!fn) return XFM_ERR_NO_DRIVER;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (inner=%llu outer=%llu ld=%llu box=%u,%u base=%p)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems, box_inner, box_outer,
              base);
    return XFM_ERR_BAD_ARG;
  }
  return 0;
}

template <int BLOCK_N, int A_MN, int B_MN>
static int launch_gemm(const xfm_gemm_params* p, const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap map_a, map_b;
  int rc;
  if (A_MN) rc = encode_2d(&map_a, p->A, p->M, p->K, p->lda, 64, BLOCK_K);
  else rc = encode_2d(&map_a, p->A, p->K, p->M, p->lda, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  if (B_MN) rc = encode_2d(&map_b, p->B, p->N, p->K, p->ldb, 64, BLOCK_K);
  else rc = encode_2d(&map_b, p->B, p->K, p->N, p->ldb, BLOCK_K, BLOCK_N);
  if (rc) return rc;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int num_tiles = g.num_m_tiles * g.num_n_tiles * g.split_k;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(map_a, map_b, g);
  count_launch();
  return (int)cudaGetLastError();
}

template <int BLOCK_N>
static int dispatch_major(const xfm_gemm_params* p, const GemmArgs& g, cudaStream_t s) {
  if (p->a_mn_major) {
    if (p->b_mn_major) return launch_gemm<BLOCK_N, 1, 1>(p, g, s);
    return launch_gemm<BLOCK_N, 1, 0>(p, g, s);
  }
  if (p->b_mn_major) return launch_gemm<BLOCK_N, 0, 1>(p, g, s);
  return launch_gemm<BLOCK_N, 0, 0>(p, g, s);
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int gemm_bf16(const xfm_gemm_params* p, cudaStream_t stream) {
  if (!p || !p->A || !p->B || !p->C || p->M <= 0 || p->N <= 0 || p->K <= 0) {
    set_error("gemm: null pointer or non-positive shape");
    return XFM_ERR_BAD_ARG;
  }
  if ((p->lda & 7) || (p->ldb & 7) || !aligned16(p->A) || !aligned16(p->B)) {
    set_error("gemm: operands need 16-byte aligned bases and leading dimensions that are multiples of 8");
    return XFM_ERR_BAD_ARG;
  }
  const int split_k = p->split_k < 1 ? 1 : p->split_k;
  if ((split_k > 1 || p->accumulate) && p->c_dtype != 1) {
    set_error("gemm: split_k / accumulate need an f32 output");
    return XFM_ERR_BAD_ARG;
  }
  if (split_k > 1 && !p->accumulate) {
    set_error("gemm: split_k > 1 needs accumulate = 1 (caller zeroes or owns C)");
    return XFM_ERR_BAD_ARG;
  }
  if (p->act == 2 && !p->aux_in) {
    set_error("gemm: act=2 (dgelu) needs aux_in");
    return XFM_ERR_BAD_ARG;
  }
  int bn = p->block_n;
  if (bn == 0) {
    // Largest tile that still yields >= ~1 wave of CTAs; small problems fall to narrower tiles.
    const int mt = (p->M + BLOCK_M - 1) / BLOCK_M;
    bn = 256;
    while (bn > 64 && (int64_t)mt * ((p->N + bn - 1) / bn) * split_k < num_sms()) bn >>= 1;
    if (p->N <= 64) bn = 64;
    else if (p->N <= 128 && bn > 128) bn = 128;
  }
  GemmArgs g;
  g.M = p->M; g.N = p->N; g.K = p->K;
  g.num_m_tiles = (p->M + BLOCK_M - 1) / BLOCK_M;
  g.num_n_tiles = (p->N + bn - 1) / bn;
  g.kb_total = (p->K + BLOCK_K - 1) / BLOCK_K;
  g.split_k = split_k > g.kb_total ? g.kb_total : split_k;
  g.c_dtype = p->c_dtype; g.accumulate = p->accumulate; g.act = p->act; g.res_dtype = p->res_dtype;
  g.rows_per_group = p->rows_per_group > 0 ? p->rows_per_group : 1;
  g.ldc = p->ldc; g.ld_aux_in = p->ld_aux_in; g.ld_aux_out = p->ld_aux_out; g.ld_res = p->ld_res;
  g.C = p->C; g.bias = p->bias; g.aux_in = (const bf16*)p->aux_in; g.aux_out = (bf16*)p->aux_out;
  g.col_scale = p->col_scale; g.row_group_scale = p->row_group_scale; g.residual = p->residual;
  g.dropout_p = p->dropout_p; g.dropout_seed = p->dropout_seed;
  const int c_al = p->c_dtype == 0 ? 7 : 3;
  bool vec = aligned16(p->C) && (p->ldc & c_al) == 0;
  if (p->aux_in) vec = vec && aligned16(p->aux_in) && (p->ld_aux_in & 7) == 0;
  if (p->aux_out) vec = vec && aligned16(p->aux_out) && (p->ld_aux_out & 7) == 0;
  if (p->residual) vec = vec && aligned16(p->residual) && (p->ld_res & (p->res_dtype == 0 ? 7 : 3)) == 0;
  g.vec_ok = vec ? 1 : 0;
  switch (bn) {
    case 64: return dispatch_major<64>(p, g, stream);
    case 128: return dispatch_major<128>(p, g, stream);
    case 256: return dispatch_major<256>(p, g, stream);
    default: set_error("gemm: block_n must be 0, 64, 128 or 256"); return XFM_ERR_BAD_ARG;
  }
}

}  // namespace xfm
