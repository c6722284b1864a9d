// K4 — text / fusion SELF-attention (xroberta.py:243-284 with the additive key mask of :966-970) on tcgen05 + TMEM + TMA.
//
// Sequences are 40 tokens, far below the 128-row MMA tile, so THREE consecutive samples are packed into one tile
// ("slot" s = rows / columns [40 s, 40 s + 40), rows 120..127 are padding): S = Q K^T is computed for the whole 128 x 128
// tile and only the block diagonal (a sample's queries against its own keys) is kept; the probabilities of the other
// blocks are written as zeros, so O = P V and the three backward products come out per sample from single MMA chains.
// Samples of a tile are consecutive rows of the activation matrix, so each operand is ONE TMA box of 120 rows.
//
// Forward: warp 0 TMA, warp 1 single-thread tcgen05.mma, two softmax warpgroups (thread = query row) that ping-pong over
// consecutive tiles.  Backward: ONE kernel computes S and dP once per tile, writes P (with the dropout mask) and dS to
// shared memory once, and issues all three products from them — dQ = dS K (dS read K-major), dK = dS^T Q and dV = P^T dO
// (the same buffers read MN-major) — so the elementwise work, which bounds this shape, is done once instead of twice.
// The dropout mask is the stateless (seed, (b, h, q, key)) hash shared with the mma.sync kernels (attention.cu::drop_keep).
#include "common.cuh"
#include "internal.h"

namespace xfm {

constexpr int ST_HD = 64;
constexpr int ST_L = 40;                 // tokens per sample
constexpr int ST_SLOTS = 3;              // samples per tile
constexpr int ST_ROWS = ST_SLOTS * ST_L; // 120 live rows
constexpr int ST_THREADS = 64 + 256;
constexpr int ST_OPB = ST_ROWS * 128;    // bytes one TMA box delivers
constexpr float ST_LOG2E = 1.4426950408889634f;

struct SAttnArgs {
  bf16* out;
  int64_t o_stride;
  float* lse;            // [B, H, 40]
  const float* kmask;    // additive [B, 40] or null
  int B, H, n_grp, n_tiles;
  float scale, dropout_p;
  uint64_t seed;
  const uint64_t* salt;
  const float* delta;    // [B, H, 40]
  bf16 *dq, *dk, *dv;
  int64_t dq_stride, dk_stride, dv_stride;
};

XFM_DEVINL void st_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
XFM_DEVINL void st_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
XFM_DEVINL void st_st_bf16x8(uint8_t* dst, const float (&p)[8]) {
  uint4 u;
  __nv_bfloat162 t0 = __floats2bfloat162_rn(p[0], p[1]), t1 = __floats2bfloat162_rn(p[2], p[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(p[4], p[5]), t3 = __floats2bfloat162_rn(p[6], p[7]);
  u.x = *(uint32_t*)&t0; u.y = *(uint32_t*)&t1; u.z = *(uint32_t*)&t2; u.w = *(uint32_t*)&t3;
  *(uint4*)dst = u;
}

// Columns a warp has to look at: the union of the diagonal blocks of its 32 rows.
XFM_DEVINL void st_union(int quad, int& ulo, int& uhi) {
  const int s_lo = (quad * 32) / ST_L;
  int s_hi = (quad * 32 + 31) / ST_L;
  s_hi = s_hi > ST_SLOTS - 1 ? ST_SLOTS - 1 : s_hi;
  ulo = s_lo * ST_L;
  uhi = s_hi * ST_L + ST_L;
}

// ================================================================================================ forward
constexpr int SF_STAGE = 2 * 3 * 16384 + 2 * 32768 + 2 * 2 * 128 * 4 + 256;   // per-warp output staging: 8 x [32 rows][128 B]
constexpr int SF_SMEM = SF_STAGE + 8 * 4096;
static_assert(SF_STAGE % 128 == 0, "staging alignment");

__global__ void __launch_bounds__(ST_THREADS, 1)
sattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const SAttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* sQ = smem;                 // [2][16384]
  uint8_t* sK = sQ + 2 * 16384;
  uint8_t* sV = sK + 2 * 16384;
  uint8_t* sP = sV + 2 * 16384;       // [2][32768]: two 64-key blocks of 128 rows
  float* smk = (float*)(sP + 2 * 32768);   // [tile buffer][unit parity][128] key mask * log2(e)
  uint64_t* bars = (uint64_t*)(smk + 2 * 2 * 128);
  uint64_t *in_full = bars, *in_empty = bars + 2, *v_full = bars + 4, *v_empty = bars + 6, *s_full = bars + 8,
           *p_full = bars + 10, *o_full = bars + 12, *o_empty = bars + 14;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pad rows, unused slots and the off-diagonal blocks of P stay zero for the whole kernel
  for (int i = threadIdx.x; i < (2 * 3 * 16384 + 2 * 32768) / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&in_full[t], 1);
      mbar_init(&in_empty[t], 1);
      mbar_init(&v_full[t], 1);
      mbar_init(&v_empty[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&o_full[t], 1);
      mbar_init(&o_empty[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t TM_O = 256;

  const int t0 = (int)((int64_t)blockIdx.x * a.n_tiles / gridDim.x);
  const int t1 = (int)((int64_t)(blockIdx.x + 1) * a.n_tiles / gridDim.x);
  const int n_units = (t1 - t0 + 1) / 2;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ph = 0;
      for (int j = 0; j < n_units; ++j) {
        for (int t = 0; t < 2; ++t) {
          const int tile = t0 + 2 * j + t;
          if (tile >= t1) break;
          const int h = tile / a.n_grp, grp = tile % a.n_grp;
          mbar_wait_relaxed(&in_empty[t], ph ^ 1);
          mbar_arrive_expect_tx(&in_full[t], 2 * ST_OPB);
          tma_load_2d(sQ + t * 16384, &map_q, &in_full[t], h * ST_HD, grp * ST_ROWS);
          tma_load_2d(sK + t * 16384, &map_k, &in_full[t], h * ST_HD, grp * ST_ROWS);
          mbar_wait_relaxed(&v_empty[t], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[t], ST_OPB);
          tma_load_2d(sV + t * 16384, &map_v, &v_full[t], h * ST_HD, grp * ST_ROWS);
        }
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, ST_HD, 0, 1);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      uint32_t ph = 0;
      for (int j = 0; j < n_units; ++j) {
        for (int t = 0; t < 2; ++t) {   // S buffer t is free: the previous unit's p_full[t] was consumed below
          if (t0 + 2 * j + t >= t1) break;
          mbar_wait(&in_full[t], ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + t * 128, make_smem_desc(aQ + t * 16384 + k * 32, 16, 1024),
                      make_smem_desc(aK + t * 16384 + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
          umma_commit(&in_empty[t]);
        }
        for (int t = 0; t < 2; ++t) {
          if (t0 + 2 * j + t >= t1) break;
          mbar_wait(&v_full[t], ph);
          mbar_wait(&p_full[t], ph);
          mbar_wait(&o_empty[t], ph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + TM_O + t * 64, make_smem_desc(aP + t * 32768 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                      make_smem_desc(aV + t * 16384 + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(&o_full[t]);
          umma_commit(&v_empty[t]);
        }
        ph ^= 1;
      }
    }
    __syncwarp();
  } else {
    const int t = (warp - 2) >> 2;             // tile buffer of this warpgroup
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int slot = r / ST_L, q = r % ST_L;
    const int blk_lo = slot * ST_L;
    int ulo, uhi;
    st_union(quad, ulo, uhi);
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_s = lane_base + (uint32_t)(t * 128), t_o = lane_base + TM_O + (uint32_t)(t * 64);
    const float scale2 = a.scale * ST_LOG2E;
    const bool drop_on = a.dropout_p > 0.f;
    const float inv_keep = drop_on ? 1.0f / (1.0f - a.dropout_p) : 1.0f;
    const uint32_t seed_mix = drop_seed_mix(a.seed + *a.salt), thr = drop_threshold(a.dropout_p);
    uint8_t* myP = sP + t * 32768 + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;
    uint32_t ph = 0;
    for (int j = 0; j < n_units; ++j) {
      const int tile = t0 + 2 * j + t;
      if (tile >= t1) break;
      const int h = tile / a.n_grp, grp = tile % a.n_grp;
      const int b = grp * ST_SLOTS + slot;
      const bool valid = slot < ST_SLOTS && b < a.B;
      float* mk = smk + (t * 2 + (j & 1)) * 128;
      {
        const int64_t flat = (int64_t)grp * ST_ROWS + r;   // == b * 40 + q for live rows
        mk[r] = (a.kmask && r < ST_ROWS && flat < (int64_t)a.B * ST_L) ? __ldg(a.kmask + flat) * ST_LOG2E : 0.f;
      }
      st_named_bar(1 + t, 128);
      const uint64_t pair_base = ((((uint64_t)(valid ? b : 0) * a.H + h) * ST_L + q) * (uint64_t)ST_L) >> 1;
      const uint32_t pb_lo = (uint32_t)pair_base, pb_hi = (uint32_t)(pair_base >> 32);
      mbar_wait(&s_full[t], ph);
      tc_fence_after();
      float m = -INFINITY;
      {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
        for (int c0 = ulo & ~31; c0 < uhi; c0 += 32) {   // warp-uniform bounds
          uint32_t v[32];
          tmem_ld_32x32(t_s + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if ((unsigned)(c0 + e - blk_lo) < (unsigned)ST_L)
              m4[e & 3] = fmaxf(m4[e & 3], fmaf(__uint_as_float(v[e]), scale2, mk[c0 + e]));
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c0 = ulo & ~31; c0 < uhi; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_s + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int col8 = c0 + g8 * 8;
          if (col8 < ulo || col8 >= uhi) continue;     // warp-uniform; columns outside the union are never written
          float p[8];
          const bool inb = valid && (unsigned)(col8 - blk_lo) < (unsigned)ST_L;   // blocks are multiples of 8 wide
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float p0 = 0.f, p1 = 0.f;
            if (inb) {
              const int c = col8 + e;
              p0 = ex2_approx(fmaf(__uint_as_float(v[g8 * 8 + e]), scale2, mk[c]) - m);
              p1 = ex2_approx(fmaf(__uint_as_float(v[g8 * 8 + e + 1]), scale2, mk[c + 1]) - m);
              s4[e & 3] += p0;
              s4[(e + 1) & 3] += p1;
              if (drop_on) {
                const uint32_t lo = pb_lo + (uint32_t)((c - blk_lo) >> 1);
                const uint32_t keep = drop_keep_pair(seed_mix, lo, pb_hi + (lo < pb_lo ? 1u : 0u), thr);
                p0 = (keep & 1u) ? p0 * inv_keep : 0.f;
                p1 = (keep & 2u) ? p1 * inv_keep : 0.f;
              }
            }
            p[e] = p0;
            p[e + 1] = p1;
          }
          st_st_bf16x8(myP + (col8 >> 6) * 16384 + ((((col8 & 63) >> 3) ^ sw) << 4), p);
        }
      }
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      mbar_wait(&o_full[t], ph);
      tc_fence_after();
      uint32_t o[2][32];
      tmem_ld_32x32(t_o, o[0]);
      tmem_ld_32x32(t_o + 32, o[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[t]);
      {
        // thread = row would store 16-byte pieces of 32 different rows per instruction; the rows of a tile are consecutive
        // rows of the output (grp * 120 + r), so each warp stages its 32 rows and stores four whole 128-byte rows per instruction
        uint8_t* stg = smem + SF_STAGE + (warp - 2) * 4096;
        const float inv = valid ? 1.0f / sum : 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            float vv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(o[hh][e + k]) * inv;
            st_st_bf16x8(stg + lane * 128 + (((hh * 4 + (e >> 3)) ^ (lane & 7)) << 4), vv);
          }
        if (valid && a.lse) a.lse[((int64_t)b * a.H + h) * ST_L + q] = (m + log2f(sum)) * 0.6931471805599453f;
        __syncwarp();
        const int ch = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int sl = it * 4 + (lane >> 3);
          const int v2 = __shfl_sync(0xffffffffu, (int)valid, sl);
          const uint4 v4 = *(const uint4*)(stg + sl * 128 + ((ch ^ (sl & 7)) << 4));
          if (v2) *(uint4*)(a.out + ((int64_t)grp * ST_ROWS + quad * 32 + sl) * a.o_stride + h * ST_HD + ch * 8) = v4;
        }
        __syncwarp();
      }
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================ backward (fused)
// Per tile: S = Q K^T and dP = dO V^T into TMEM stage i & 1; the two warpgroups split each warp's column union in halves and
// write P~ = keep * P / (1-p) and dS = P o (keep * dP / (1-p) - delta) to shared memory; then dQ = dS K, dK = dS^T Q and
// dV = P~^T dO accumulate into the TMEM columns S / dP occupied.  Operands and TMEM are double-buffered so the next tile's
// S / dP are ready when the elementwise stage gets there; P~ / dS are single-buffered.
constexpr int SB_STAGE = 2 * 4 * 16384 + 2 * 32768 + 2 * 128 * 4 + 256;   // per-warp output staging: 8 x [32 rows][64 B]
constexpr int SB_SMEM = SB_STAGE + 8 * 2048;
static_assert(SB_STAGE % 128 == 0, "staging alignment");

__global__ void __launch_bounds__(ST_THREADS, 1)
sattn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const SAttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* sQ = smem;                  // [2][16384] each
  uint8_t* sdO = sQ + 2 * 16384;
  uint8_t* sK = sdO + 2 * 16384;
  uint8_t* sV = sK + 2 * 16384;
  uint8_t* sP = sV + 2 * 16384;        // 32768
  uint8_t* sdS = sP + 32768;           // 32768
  float* smk = (float*)(sdS + 32768);  // [2][128]
  uint64_t* bars = (uint64_t*)(smk + 2 * 128);
  uint64_t *in_full = bars, *in_empty = bars + 2, *sd_full = bars + 4, *tm_empty = bars + 6, *out_full = bars + 8,
           *ds_full = bars + 10;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (2 * 4 * 16384 + 2 * 32768) / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&in_full[t], 1);
      mbar_init(&in_empty[t], 1);
      mbar_init(&sd_full[t], 1);
      mbar_init(&tm_empty[t], 8);
      mbar_init(&out_full[t], 1);
    }
    mbar_init(ds_full, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int t0 = (int)((int64_t)blockIdx.x * a.n_tiles / gridDim.x);
  const int t1 = (int)((int64_t)(blockIdx.x + 1) * a.n_tiles / gridDim.x);
  const int n = t1 - t0;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n; ++i) {
        const int st = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        const int tile = t0 + i;
        const int h = tile / a.n_grp, grp = tile % a.n_grp;
        mbar_wait_relaxed(&in_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&in_full[st], 4 * ST_OPB);
        tma_load_2d(sQ + st * 16384, &map_q, &in_full[st], h * ST_HD, grp * ST_ROWS);
        tma_load_2d(sK + st * 16384, &map_k, &in_full[st], h * ST_HD, grp * ST_ROWS);
        tma_load_2d(sdO + st * 16384, &map_do, &in_full[st], h * ST_HD, grp * ST_ROWS);
        tma_load_2d(sV + st * 16384, &map_v, &in_full[st], h * ST_HD, grp * ST_ROWS);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, ST_HD, 0, 1);
      constexpr uint32_t idesc_dkv = make_idesc_bf16(128, ST_HD, 1, 1);
      const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP),
                     adS = smem_u32(sdS);
      auto issue_sdp = [&](int i) {
        const int st = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        mbar_wait(&in_full[st], par);
        mbar_wait(&tm_empty[st], par ^ 1);
        tc_fence_after();
        const uint32_t tm = tmem_base + st * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm, make_smem_desc(aQ + st * 16384 + k * 32, 16, 1024), make_smem_desc(aK + st * 16384 + k * 32, 16, 1024),
                    idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm + 128, make_smem_desc(adO + st * 16384 + k * 32, 16, 1024),
                    make_smem_desc(aV + st * 16384 + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&sd_full[st]);
      };
      if (n > 0) issue_sdp(0);
      for (int i = 0; i < n; ++i) {
        const int st = i & 1;
        if (i + 1 < n) issue_sdp(i + 1);
        mbar_wait(ds_full, (uint32_t)i & 1u);
        tc_fence_after();
        const uint32_t tm = tmem_base + st * 256;
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dQ = dS K: dS K-major, K MN-major
          umma_bf16(tm, make_smem_desc(adS + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    make_smem_desc(aK + st * 16384 + k * 2048, 8192, 1024), idesc_dq, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dK = dS^T Q: the same dS buffer read MN-major (64-key blocks 16 KB apart)
          umma_bf16(tm + 64, make_smem_desc(adS + k * 2048, 16384, 1024), make_smem_desc(aQ + st * 16384 + k * 2048, 8192, 1024),
                    idesc_dkv, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV = P~^T dO
          umma_bf16(tm + 128, make_smem_desc(aP + k * 2048, 16384, 1024), make_smem_desc(adO + st * 16384 + k * 2048, 8192, 1024),
                    idesc_dkv, k > 0 ? 1u : 0u);
        umma_commit(&out_full[st]);
        umma_commit(&in_empty[st]);
      }
    }
    __syncwarp();
  } else {
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int slot = r / ST_L, q = r % ST_L;
    const int blk_lo = slot * ST_L;
    int ulo, uhi;
    st_union(quad, ulo, uhi);
    const int umid = ulo + (((uhi - ulo) / 2 + 7) & ~7);
    const int cb = wg == 0 ? ulo : umid, ce = wg == 0 ? umid : uhi;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float scale2 = a.scale * ST_LOG2E;
    const bool drop_on = a.dropout_p > 0.f;
    const float inv_keep = drop_on ? 1.0f / (1.0f - a.dropout_p) : 1.0f;
    const uint32_t seed_mix = drop_seed_mix(a.seed + *a.salt), thr = drop_threshold(a.dropout_p);
    uint8_t* myP = sP + (r >> 3) * 1024 + (r & 7) * 128;
    uint8_t* myD = sdS + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;
    for (int i = 0; i < n; ++i) {
      const int st = i & 1;
      const uint32_t par = (uint32_t)(i >> 1) & 1u;
      const int tile = t0 + i;
      const int h = tile / a.n_grp, grp = tile % a.n_grp;
      const int b = grp * ST_SLOTS + slot;
      const bool valid = slot < ST_SLOTS && b < a.B;
      float* mk = smk + st * 128;
      if (wg == 0) {
        const int64_t flat = (int64_t)grp * ST_ROWS + r;
        mk[r] = (a.kmask && r < ST_ROWS && flat < (int64_t)a.B * ST_L) ? __ldg(a.kmask + flat) * ST_LOG2E : 0.f;
      }
      st_named_bar(1, 256);
      const int64_t st_row = ((int64_t)(valid ? b : 0) * a.H + h) * ST_L + q;
      const float lse2 = valid ? __ldg(a.lse + st_row) * ST_LOG2E : 0.f;
      const float dl = valid ? __ldg(a.delta + st_row) : 0.f;
      const uint64_t pair_base = ((uint64_t)st_row * (uint64_t)ST_L) >> 1;
      const uint32_t pb_lo = (uint32_t)pair_base, pb_hi = (uint32_t)(pair_base >> 32);
      const uint32_t tm = lane_base + (uint32_t)(st * 256);
      mbar_wait(&sd_full[st], par);
      tc_fence_after();
      for (int c0 = cb; c0 < ce; c0 += 16) {
        uint32_t vs[16], vp[16];
        st_ld16(tm + c0, vs);
        st_ld16(tm + 128 + c0, vp);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          const int col8 = c0 + g8 * 8;
          if (col8 >= ce) continue;     // warp-uniform
          const bool inb = valid && (unsigned)(col8 - blk_lo) < (unsigned)ST_L;
          float pp[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
            if (inb) {
              const int c = col8 + e;
              p0 = ex2_approx(fmaf(__uint_as_float(vs[g8 * 8 + e]), scale2, mk[c] - lse2));
              p1 = ex2_approx(fmaf(__uint_as_float(vs[g8 * 8 + e + 1]), scale2, mk[c + 1] - lse2));
              float dp0 = __uint_as_float(vp[g8 * 8 + e]), dp1 = __uint_as_float(vp[g8 * 8 + e + 1]);
              float k0 = 1.f, k1 = 1.f;
              if (drop_on) {
                const uint32_t lo = pb_lo + (uint32_t)((c - blk_lo) >> 1);
                const uint32_t keep = drop_keep_pair(seed_mix, lo, pb_hi + (lo < pb_lo ? 1u : 0u), thr);
                k0 = (keep & 1u) ? inv_keep : 0.f;
                k1 = (keep & 2u) ? inv_keep : 0.f;
              }
              d0 = p0 * (dp0 * k0 - dl);
              d1 = p1 * (dp1 * k1 - dl);
              p0 *= k0;
              p1 *= k1;
            }
            pp[e] = p0;
            pp[e + 1] = p1;
            ds[e] = d0;
            ds[e + 1] = d1;
          }
          const int off8 = (col8 >> 6) * 16384 + ((((col8 & 63) >> 3) ^ sw) << 4);
          st_st_bf16x8(myP + off8, pp);
          st_st_bf16x8(myD + off8, ds);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      mbar_wait(&out_full[st], par);
      tc_fence_after();
      // TMEM columns of the stage: dQ [0,64)  dK [64,128)  dV [128,192).  Warpgroup 0 drains [0,96), warpgroup 1 [96,192).
      uint32_t o[3][32];
      const uint32_t src = tm + (uint32_t)(wg * 96);
      tmem_ld_32x32(src, o[0]);
      tmem_ld_32x32(src + 32, o[1]);
      tmem_ld_32x32(src + 64, o[2]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tm_empty[st]);
      {
        // wg 0: dQ[0:32) dQ[32:64) dK[0:32)     wg 1: dK[32:64) dV[0:32) dV[32:64).  Each 32-column part is staged per warp
        // ([32 rows][64 B], chunk index xor-ed with (row >> 1) & 3: conflict-free both ways) and stored eight rows x 64 bytes
        // per instruction: whole sectors instead of 16-byte pieces of 32 rows.  Tile rows are consecutive global rows.
        uint8_t* stg = smem + SB_STAGE + (warp - 2) * 2048;
        const int64_t row0 = (int64_t)grp * ST_ROWS + quad * 32;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          bf16* dst;
          int64_t ld;
          float mul;
          if (wg == 0) {
            dst = part < 2 ? a.dq + part * 32 : a.dk;
            ld = part < 2 ? a.dq_stride : a.dk_stride;
            mul = a.scale;
          } else {
            dst = part == 0 ? a.dk + 32 : a.dv + (part - 1) * 32;
            ld = part == 0 ? a.dk_stride : a.dv_stride;
            mul = part == 0 ? a.scale : 1.0f;
          }
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            float vv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(o[part][e + k]) * mul;
            st_st_bf16x8(stg + lane * 64 + ((((e >> 3)) ^ ((lane >> 1) & 3)) << 4), vv);
          }
          __syncwarp();
          const int ch = lane & 3;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int sl = it * 8 + (lane >> 2);
            const int v2 = __shfl_sync(0xffffffffu, (int)valid, sl);
            const uint4 v4 = *(const uint4*)(stg + sl * 64 + ((ch ^ ((sl >> 1) & 3)) << 4));
            if (v2) *(uint4*)(dst + (row0 + sl) * ld + h * ST_HD + ch * 8) = v4;
          }
          __syncwarp();
        }
      }
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
static int st_encode_rows(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems) {
  auto fn = get_tensor_map_encoder();
  if (!fn) return XFM_ERR_NO_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {ST_HD, ST_ROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("self attention: cuTensorMapEncodeTiled failed: %d", (int)r);
    return XFM_ERR_BAD_ARG;
  }
  return 0;
}

// Instantiated shape: 40-token text / fusion sequences (BASELINE configs[1] max_tokens), optional additive key mask,
// optional dropout; every sample attends to its own keys.
bool self_attention_tc_supported(const xfm_attn_params* p) {
  return p->head_dim == ST_HD && p->Lq == ST_L && p->Lk == ST_L && !p->bias && !p->kv_index && !p->rel_table &&
         (p->Bkv == 0 || p->Bkv == p->B) && ((uintptr_t)p->q & 15) == 0 && ((uintptr_t)p->k & 15) == 0 &&
         ((uintptr_t)p->v & 15) == 0 && ((p->q_stride | p->k_stride | p->v_stride | p->o_stride) & 7) == 0;
}

static int st_fill(const xfm_attn_params* p, SAttnArgs& a) {
  a.out = (bf16*)p->out; a.o_stride = p->o_stride; a.lse = p->lse; a.kmask = p->kmask;
  a.B = p->B; a.H = p->H; a.scale = p->scale; a.dropout_p = p->dropout_p; a.seed = p->dropout_seed; a.salt = seed_salt_ptr();
  a.n_grp = (p->B + ST_SLOTS - 1) / ST_SLOTS;
  a.n_tiles = a.n_grp * a.H;
  a.delta = p->delta;
  a.dq = (bf16*)p->dq; a.dk = (bf16*)p->dk; a.dv = (bf16*)p->dv;
  a.dq_stride = p->dq_stride; a.dk_stride = p->dk_stride; a.dv_stride = p->dv_stride;
  return a.n_tiles < num_sms() ? a.n_tiles : num_sms();
}

int self_attention_fwd_tc(const xfm_attn_params* p, cudaStream_t s) {
  SAttnArgs a;
  const int grid = st_fill(p, a);
  if (grid <= 0) return 0;
  const uint64_t cols = (uint64_t)a.H * ST_HD, rows = (uint64_t)a.B * ST_L;
  CUtensorMap mq, mk, mv;
  int rc = st_encode_rows(&mq, p->q, cols, rows, p->q_stride);
  if (!rc) rc = st_encode_rows(&mk, p->k, cols, rows, p->k_stride);
  if (!rc) rc = st_encode_rows(&mv, p->v, cols, rows, p->v_stride);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(sattn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  sattn_fwd_tc_kernel<<<grid, ST_THREADS, SF_SMEM, s>>>(mq, mk, mv, a);
  count_launch();
  return (int)cudaGetLastError();
}

int self_attention_bwd_tc(const xfm_attn_params* p, cudaStream_t s) {
  if ((((uintptr_t)p->dout | (uintptr_t)p->dq | (uintptr_t)p->dk | (uintptr_t)p->dv) & 15) ||
      ((p->do_stride | p->dq_stride | p->dk_stride | p->dv_stride) & 7)) {
    set_error("self attention bwd: operands must be 16-byte aligned with row strides that are multiples of 8");
    return XFM_ERR_BAD_ARG;
  }
  SAttnArgs a;
  const int grid = st_fill(p, a);
  if (grid <= 0) return 0;
  const uint64_t cols = (uint64_t)a.H * ST_HD, rows = (uint64_t)a.B * ST_L;
  CUtensorMap mq, mdo, mk, mv;
  int rc = st_encode_rows(&mq, p->q, cols, rows, p->q_stride);
  if (!rc) rc = st_encode_rows(&mdo, p->dout, cols, rows, p->do_stride);
  if (!rc) rc = st_encode_rows(&mk, p->k, cols, rows, p->k_stride);
  if (!rc) rc = st_encode_rows(&mv, p->v, cols, rows, p->v_stride);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(sattn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  sattn_bwd_tc_kernel<<<grid, ST_THREADS, SB_SMEM, s>>>(mq, mdo, mk, mv, a);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace xfm
