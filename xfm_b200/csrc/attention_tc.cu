// K2 — BEiT / ViT self-attention forward on tcgen05 + TMEM + TMA (head_dim 64, Lq = Lk <= 208 tokens: one pass, no
// online softmax).  Reference arithmetic: beit2.py:126-166  (q*d^-1/2) k^T + relative_position_bias -> softmax -> @ v.
//
// One persistent CTA per SM walks a contiguous range of (head, sample) items.  Per item:
//   warp 0      TMA: Q tile (256 rows), K (Lpad rows), V (Lpad rows) boxes straight out of the fused qkv activation
//               (row stride 3*D) into SWIZZLE_128B shared-memory tiles.
//   warp 1      one thread issues tcgen05.mma:  S_t = Q_t K^T  (M=128, N=Lpad, K=64) for both 128-row query tiles into
//               TMEM, later O_t = P_t V (M=128, N=64, K=Lpad) which overwrites the first 64 columns of S_t.
//   warps 2..9  two softmax warpgroups, one per query tile, thread = query row: pass 1 pulls S out of TMEM, applies
//               scale + relative-position bias (gathered from the head's 732-entry table held in shared memory through the
//               closed form of beit2.py:104-114's index, so the [H, N, N] bias tensor is never read), writes the logits
//               back to TMEM and tracks the row maximum; pass 2 re-reads them, exponentiates (ex2), accumulates the row
//               sum and stores bf16 P as the K-major A operand of the second MMA; finally O / rowsum -> global, lse.
// Scores, probabilities and the bias tensor never touch HBM.
#include "common.cuh"
#include "internal.h"

namespace xfm {

constexpr int TC_HD = 64;
constexpr int TC_THREADS = 64 + 256;
constexpr int TC_QROWS = 256;

struct VitAttnArgs {
  bf16* out;
  int64_t o_stride;
  float* lse;            // [B, H, L] natural-log sum-exp of the scaled + biased logits
  const float* table;    // [T, H] relative_position_bias_table or null
  int B, H, L, Lpad, W, T, ntiles;
  float scale;
  int items_per_cta;
};

XFM_DEVINL void tmem_ld_32x32_16(uint32_t taddr, uint32_t (&r)[32]) {  // 16 columns into r[0..15]
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
XFM_DEVINL void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
XFM_DEVINL void tmem_st_32x32_16(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
XFM_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
XFM_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1)
vit_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v, const VitAttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kv_bytes = a.Lpad * 128;
  const int nkb = (a.Lpad + 63) / 64;          // 64-key blocks of the P operand
  const int p_bytes = nkb * 16384;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TC_QROWS * 128;
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sP = sV + kv_bytes;                  // [2][p_bytes]
  float* tb = (float*)(sP + 2 * p_bytes);       // [T] bias table of the current head, pre-multiplied by log2(e)
  int* off = (int*)(tb + ((a.T + 31) & ~31));   // [256] column part of the relative-position index
  uint64_t* bars = (uint64_t*)(off + 256);
  uint64_t *qk_full = bars, *qk_empty = bars + 1, *v_full = bars + 2, *v_empty = bars + 3;
  uint64_t *s_full = bars + 4, *s_empty = bars + 6, *p_full = bars + 8, *o_full = bars + 10;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_empty[t], 4);
      mbar_init(&p_full[t], 4);
      mbar_init(&o_full[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  for (int j = threadIdx.x; j < 256; j += blockDim.x) {
    const int jj = min(j, a.L - 1);  // padded key columns reuse the last valid entry (their logits are forced to -inf)
    off[j] = jj >= 1 ? ((jj - 1) / a.W) * (2 * a.W - 1) + (jj - 1) % a.W : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_items = a.B * a.H;
  const int item0 = blockIdx.x * a.items_per_cta;
  const int item1 = min(n_items, item0 + a.items_per_cta);

  if (warp == 0) {
    if (lane == 0) {
      for (int it = item0; it < item1; ++it) {
        const uint32_t ph = (uint32_t)(it - item0) & 1u;
        const int h = it / a.B, b = it % a.B;
        mbar_wait_relaxed(qk_empty, ph ^ 1);
        mbar_arrive_expect_tx(qk_full, TC_QROWS * 128 + kv_bytes);
        tma_load_2d(sQ, &map_q, qk_full, h * TC_HD, b * a.L);
        tma_load_2d(sK, &map_k, qk_full, h * TC_HD, b * a.L);
        mbar_wait_relaxed(v_empty, ph ^ 1);
        mbar_arrive_expect_tx(v_full, kv_bytes);
        tma_load_2d(sV, &map_v, v_full, h * TC_HD, b * a.L);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, a.Lpad, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, TC_HD, 0, 1);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      for (int it = item0; it < item1; ++it) {
        const uint32_t ph = (uint32_t)(it - item0) & 1u;
        mbar_wait(qk_full, ph);
        tc_fence_after();
        for (int t = 0; t < a.ntiles; ++t) {
          mbar_wait(&s_empty[t], ph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + t * a.Lpad, make_smem_desc(aQ + t * 16384 + k * 32, 16, 1024),
                      make_smem_desc(aK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
        }
        umma_commit(qk_empty);
        mbar_wait(v_full, ph);
        tc_fence_after();
        for (int t = 0; t < a.ntiles; ++t) {
          mbar_wait(&p_full[t], ph);
          tc_fence_after();
          const int nk = a.Lpad / 16;
          for (int k = 0; k < nk; ++k)
            umma_bf16(tmem_base + t * a.Lpad, make_smem_desc(aP + t * p_bytes + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                      make_smem_desc(aV + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(&o_full[t]);
        }
        umma_commit(v_empty);
      }
    }
    __syncwarp();
  } else {
    const int t = (warp - 2) >> 2;             // query tile of this warpgroup
    const int quad = warp & 3;                 // TMEM lane quadrant
    const int r = quad * 32 + lane;            // row inside the tile
    const int qi = t * 128 + r;                // query index inside the sample
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * a.Lpad);
    const int wgt = threadIdx.x - 64;          // 0..255 among the softmax threads
    // closed form of beit2.py:104-114: idx(i, j) = base_i - mul_i * off_j for j >= 1; column 0 and row 0 are constants
    const int qc = min(qi, a.L - 1);
    const int pi = qc - 1;
    const bool has_tab = a.table != nullptr;  // no table (VQ-KD tokenizer): every index points at tb[0] = 0
    const int base_i = !has_tab ? 0 : (qc >= 1 ? ((pi / a.W) + a.W - 1) * (2 * a.W - 1) + (pi % a.W) + a.W - 1 : a.T - 3);
    const int mul_i = (has_tab && qc >= 1) ? 1 : 0;
    const int idx_col0 = !has_tab ? 0 : (qc >= 1 ? a.T - 2 : a.T - 1);
    const float scale2 = a.scale * 1.4426950408889634f;
    uint8_t* myP = sP + t * p_bytes + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;
    int cur_h = -1;
    for (int it = item0; it < item1; ++it) {
      const uint32_t ph = (uint32_t)(it - item0) & 1u;
      const int h = it / a.B, b = it % a.B;
      if (h != cur_h) {  // (re)load this head's bias column; both warpgroups take the same branch for the same item
        named_bar_sync(1, 256);
        for (int i = wgt; i < a.T; i += 256) tb[i] = a.table ? __ldg(a.table + (int64_t)i * a.H + h) * 1.4426950408889634f : 0.f;
        named_bar_sync(1, 256);
        cur_h = h;
      }
      if (t >= a.ntiles) continue;
      mbar_wait(&s_full[t], ph);
      tc_fence_after();
      // ---- pass 1: logits (log2 domain) back into TMEM, row maximum
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      for (int c0 = 0; c0 < a.Lpad; c0 += 32) {
        uint32_t v[32];
        const bool half = c0 + 32 > a.Lpad;  // trailing 16-column chunk
        if (half) tmem_ld_32x32_16(taddr + c0, v);
        else tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (half && e >= 16) break;
          const int j = c0 + e;
          int idx = base_i - mul_i * off[j];
          if (e == 0 && c0 == 0) idx = idx_col0;
          float l = fmaf(__uint_as_float(v[e]), scale2, tb[idx]);
          if (j >= a.L) l = -INFINITY;
          m4[e & 3] = fmaxf(m4[e & 3], l);
          v[e] = __float_as_uint(l);
        }
        if (half) tmem_st_32x32_16(taddr + c0, v);
        else tmem_st_32x32(taddr + c0, v);
      }
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      tmem_st_wait();
      // ---- pass 2: p = 2^(l - m), row sum, bf16 P -> shared memory (K-major, SWIZZLE_128B)
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c0 = 0; c0 < a.Lpad; c0 += 32) {
        uint32_t v[32];
        const bool half = c0 + 32 > a.Lpad;
        if (half) tmem_ld_32x32_16(taddr + c0, v);
        else tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        uint8_t* blk = myP + (c0 >> 6) * 16384;
        const int ch0 = (c0 & 63) >> 3;  // first 16-byte chunk of this 32-column group inside the 64-key block
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          if (half && g8 >= 2) break;
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            p[e] = ex2_approx(__uint_as_float(v[g8 * 8 + e]) - m);
            s4[e & 3] += p[e];
          }
          uint4 u;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(p[0], p[1]), t1 = __floats2bfloat162_rn(p[2], p[3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(p[4], p[5]), t3 = __floats2bfloat162_rn(p[6], p[7]);
          u.x = *(uint32_t*)&t0; u.y = *(uint32_t*)&t1; u.z = *(uint32_t*)&t2; u.w = *(uint32_t*)&t3;
          *(uint4*)(blk + (((ch0 + g8) ^ sw) << 4)) = u;
        }
      }
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      // ---- epilogue: O / rowsum -> global, lse
      mbar_wait(&o_full[t], ph);
      tc_fence_after();
      const float inv = 1.0f / sum;
      const bool valid = qi < a.L;
      bf16* orow = a.out + ((int64_t)b * a.L + qi) * a.o_stride + h * TC_HD;
#pragma unroll
      for (int c0 = 0; c0 < TC_HD; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 u;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv);
            u.x = *(uint32_t*)&t0; u.y = *(uint32_t*)&t1; u.z = *(uint32_t*)&t2; u.w = *(uint32_t*)&t3;
            *(uint4*)(orow + c0 + e) = u;
          }
        }
      }
      if (valid && a.lse) a.lse[((int64_t)b * a.H + h) * a.L + qi] = (m + log2f(sum)) * 0.6931471805599453f;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[t]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ host
static int encode_rows(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_rows) {
  auto fn = get_tensor_map_encoder();
  if (!fn) return XFM_ERR_NO_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {TC_HD, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("vit attention: cuTensorMapEncodeTiled failed: %d", (int)r);
    return XFM_ERR_BAD_ARG;
  }
  return 0;
}

bool vit_attention_tc_supported(const xfm_attn_params* p) {
  return p->head_dim == TC_HD && p->Lq == p->Lk && p->Lk <= 208 && p->Lk > 16 && !p->kmask && !p->kv_index &&
         (!p->bias || p->rel_table) && !(p->dropout_p > 0.f) && (p->Bkv == 0 || p->Bkv == p->B) && (p->rel_table == nullptr || p->rel_window > 0) &&
         ((uintptr_t)p->q & 15) == 0 && ((uintptr_t)p->k & 15) == 0 && ((uintptr_t)p->v & 15) == 0 &&
         ((p->q_stride | p->k_stride | p->v_stride | p->o_stride) & 7) == 0;
}

int vit_attention_fwd_tc(const xfm_attn_params* p, cudaStream_t s) {
  VitAttnArgs a;
  a.out = (bf16*)p->out; a.o_stride = p->o_stride; a.lse = p->lse; a.table = p->rel_table;
  a.B = p->B; a.H = p->H; a.L = p->Lk; a.Lpad = (p->Lk + 15) / 16 * 16;
  a.W = p->rel_table ? p->rel_window : 1;
  a.T = p->rel_table ? (2 * a.W - 1) * (2 * a.W - 1) + 3 : 4;
  a.ntiles = (a.L + 127) / 128;
  a.scale = p->scale;
  if (p->rel_table && a.W * a.W + 1 != a.L) {
    set_error("vit attention: window %d does not match %d tokens", a.W, a.L);
    return XFM_ERR_BAD_ARG;
  }
  const uint64_t rows = (uint64_t)a.B * a.L, cols = (uint64_t)a.H * TC_HD;
  CUtensorMap mq, mk, mv;
  int rc = encode_rows(&mq, p->q, cols, rows, p->q_stride, TC_QROWS);
  if (!rc) rc = encode_rows(&mk, p->k, cols, rows, p->k_stride, a.Lpad);
  if (!rc) rc = encode_rows(&mv, p->v, cols, rows, p->v_stride, a.Lpad);
  if (rc) return rc;
  const int nkb = (a.Lpad + 63) / 64;
  const size_t smem = 1024 + (size_t)TC_QROWS * 128 + 2 * (size_t)a.Lpad * 128 + 2 * (size_t)nkb * 16384 +
                      (size_t)((a.T + 31) & ~31) * 4 + 256 * 4 + 128;
  if (smem > 227 * 1024) {
    set_error("vit attention: %zu bytes of shared memory needed", smem);
    return XFM_ERR_BAD_ARG;
  }
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(vit_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_smem = smem;
  }
  const int n_items = a.B * a.H;
  const int ctas = n_items < num_sms() ? n_items : num_sms();
  a.items_per_cta = (n_items + ctas - 1) / ctas;
  const int grid = (n_items + a.items_per_cta - 1) / a.items_per_cta;
  vit_attn_fwd_tc_kernel<<<grid, TC_THREADS, smem, s>>>(mq, mk, mv, a);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace xfm
