// HBM-bound kernels of the XFM hot path: LayerNorm fwd/bwd, LayerScale backward, column sums (bias
// gradients), embeddings, patch im2col / token assembly, mean-pool, casts, row gather / scatter.
// All use 128-bit coalesced accesses and warp-shuffle reductions; one warp owns one row of D features.
#include "common.cuh"
#include "internal.h"

namespace xfm {

// ---- dtype-generic 4-element accessors (dtype: 0 = bf16, 1 = f32); idx is an ELEMENT index, multiple of 4
XFM_DEVINL float4 ld4(const void* p, int dtype, size_t idx) {
  if (dtype == 1) return *(const float4*)((const float*)p + idx);
  uint2 u = *(const uint2*)((const bf16*)p + idx);
  float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&u.x);
  float2 b = __bfloat1622float2(*(const __nv_bfloat162*)&u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
XFM_DEVINL void st4(void* p, int dtype, size_t idx, float4 v) {
  if (dtype == 1) {
    *(float4*)((float*)p + idx) = v;
  } else {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *(uint32_t*)&a;
    u.y = *(uint32_t*)&b;
    *(uint2*)((bf16*)p + idx) = u;
  }
}

constexpr int LN_MAX_VEC = 16;  // per-lane float4 slots: D <= 32 * 4 * 16 = 2048
constexpr int LN_WARPS = 8;

// -------------------------------------------------------------------------------- LayerNorm forward
// y = (x - mean) * rstd * w + b.  Reference: torch layer_norm at beit2.py:201-205, xroberta.py:135,303,384,1328.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ w, const float* __restrict__ b,
                     void* __restrict__ y, int y_dtype, float* __restrict__ y2_f32, float* __restrict__ stats, int M,
                     int D, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + warp;
  if (row >= M) return;
  const size_t base = (size_t)row * D;
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < D) {
      v[i] = ld4(x, x_dtype, base + c);
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < D) {
      const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq += a * a + bb * bb + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < D) {
      const float4 ww = *(const float4*)(w + c), bb = *(const float4*)(b + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * ww.x + bb.x;
      o.y = (v[i].y - mean) * rstd * ww.y + bb.y;
      o.z = (v[i].z - mean) * rstd * ww.z + bb.z;
      o.w = (v[i].w - mean) * rstd * ww.w + bb.w;
      st4(y, y_dtype, base + c, o);
      if (y2_f32) *(float4*)(y2_f32 + base + c) = o;
    }
  }
  if (lane == 0 && stats) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
}

// -------------------------------------------------------------------------------- LayerNorm backward
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w;  dw += sum dy * xhat;  db += sum dy.
// Optional add_in (residual-path gradient) is added to dx.  dw / db are accumulated with one fp32 atomic per
// column per CTA (each CTA first reduces its rows in registers + shared memory).
//
// One warp owns one row at a time, but its inputs (dy, x, add_in, the row's mean / rstd) are staged through a per-warp
// two-slot shared-memory ring with cp.async: the loads of rows k+1 and k+2 are in flight while row k is reduced, at no
// register cost (the dgamma / dbeta accumulators already take 48 registers).  Loading straight into registers left one
// exposed DRAM round trip (two with add_in) per row and 17 % of DRAM bandwidth (ncu, profiles/r01_ln_bwd_*).
XFM_DEVINL void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
XFM_DEVINL void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
XFM_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
XFM_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

XFM_DEVINL float4 lds4(const uint8_t* slot, int dtype, int c) {  // 4 elements at column c of a staged row
  if (dtype == 1) return *(const float4*)(slot + (size_t)c * 4);
  const uint2 u = *(const uint2*)(slot + (size_t)c * 2);
  const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&u.x), b = __bfloat1622float2(*(const __nv_bfloat162*)&u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// DENSE: also emit what the Linear in front of this LayerNorm needs for ITS backward (xroberta.py:300-304,381-385:
// LayerNorm(dropout(dense(.)) + residual)): dx16 = bf16(dropout_mask * dx / (1 - p)) — the dgrad / wgrad operand, with the
// forward's mask re-derived from (seed, row * D + col) like the GEMM epilogue that applied it — and dbias += its column
// sums.  Replaces a dropout / cast pass and a column-sum pass over the same rows (2 launches per site, 60 sites per step).
//
// Round 2 (ncu, profiles/r02j_ln_summary.txt): with 8 warps of 162 registers the kernel was bound by instruction latency,
// not by DRAM (2 warps per scheduler, issue slots 30 % busy, 41 % of DRAM bandwidth).  Now 16 warps per CTA: the row's
// g / xhat are not kept in registers between the two passes but recomputed from the staged row (it sits in shared memory
// anyway), which brings the kernel under the 128 registers a 512-thread CTA may use; gamma is staged in shared memory too.
constexpr int LNB_WARPS = 16;

// DT >= 0 bakes the dtypes and D == NV * 128 into the instantiation (bit 0 dy fp32, bit 1 x fp32, bits 2-3 add_in: 0 none /
// 1 bf16 / 2 fp32, bit 4 dx fp32): the per-load dtype branches and column bounds checks were ~40 % of the issued
// instructions of the run-time version (DT = -1), which stays for every other shape.
template <int NV, bool DENSE, int DT>
__global__ void __launch_bounds__(LNB_WARPS * 32, 1)
layernorm_bwd_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ x, int x_dtype,
                     const float* __restrict__ stats, const float* __restrict__ w, const void* __restrict__ add_in,
                     int add_dtype, void* __restrict__ dx, int dx_dtype, float* __restrict__ dw, float* __restrict__ db,
                     int M, int D, int rows_per_cta, int nw,   // nw: warps that own a staging ring (<= LNB_WARPS)
                     int red_rows,                             // warps per group of the final reduction (divides LNB_WARPS)
                     bf16* __restrict__ dx16, float* __restrict__ dbias, float drop_p, uint64_t drop_seed,
                     const uint64_t* __restrict__ salt) {
  extern __shared__ __align__(16) uint8_t ln_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool FIX = DT >= 0;
  if (FIX) {
    D = NV * 128;
    dy_dtype = DT & 1;
    x_dtype = (DT >> 1) & 1;
    add_dtype = ((DT >> 2) & 3) == 2;
    dx_dtype = (DT >> 4) & 1;
    if (((DT >> 2) & 3) == 0) add_in = nullptr;
  }
  const int dy_b = D * (dy_dtype == 1 ? 4 : 2), x_b = D * (x_dtype == 1 ? 4 : 2);
  const int add_b = add_in ? D * (add_dtype == 1 ? 4 : 2) : 0;
  const int slot_b = dy_b + x_b + add_b + 16;   // + the row's (mean, rstd)
  const float* w_s = (const float*)ln_smem;     // gamma, staged once per CTA
  uint8_t* ring = ln_smem + (size_t)D * 4 + (size_t)warp * 2 * slot_b;
  const uint32_t ring_a = smem_u32(ring);
  float4 aw[NV], ab[NV], ad[DENSE ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) aw[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (DENSE ? NV : 1); ++i) ad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float inv_keep = (DENSE && drop_p > 0.f) ? 1.0f / (1.0f - drop_p) : 1.0f;
  const uint32_t seed_mix = drop_seed_mix(drop_seed + *salt), thr = drop_threshold(drop_p);
  const int row0 = blockIdx.x * rows_per_cta;
  const int row1 = (warp < nw) ? min(M, row0 + rows_per_cta) : 0;   // ring-less warps only join the final reduction

  auto issue = [&](int row, int slot) {   // stage one row; always commits a group so the wait counts stay uniform
    if (row < row1) {
      const uint32_t dst = ring_a + slot * slot_b;
      const uint8_t* s_dy = (const uint8_t*)dy + (size_t)row * dy_b;
      const uint8_t* s_x = (const uint8_t*)x + (size_t)row * x_b;
      for (int o = lane * 16; o < dy_b; o += 512) cp_async16(dst + o, s_dy + o);
      for (int o = lane * 16; o < x_b; o += 512) cp_async16(dst + dy_b + o, s_x + o);
      if (add_in) {
        const uint8_t* s_a = (const uint8_t*)add_in + (size_t)row * add_b;
        for (int o = lane * 16; o < add_b; o += 512) cp_async16(dst + dy_b + x_b + o, s_a + o);
      }
      if (lane == 0) cp_async8(dst + dy_b + x_b + add_b, stats + 2 * (size_t)row);
    }
    cp_async_commit();
  };

  int row = row0 + warp;
  issue(row, 0);
  issue(row + nw, 1);
  for (int c = threadIdx.x * 4; c < D; c += LNB_WARPS * 32 * 4) *(float4*)(ln_smem + (size_t)c * 4) = *(const float4*)(w + c);
  __syncthreads();
  for (int k = 0; row < row1; ++k, row += nw) {
    cp_async_wait<1>();
    __syncwarp();
    const uint8_t* slot = ring + (k & 1) * slot_b;
    const float2 mr = *(const float2*)(slot + dy_b + x_b + add_b);
    const float mean = mr.x, rstd = mr.y;
    const size_t base = (size_t)row * D;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {   // pass 1: the row's two means, dgamma / dbeta partial sums
      const int c = (i * 32 + lane) * 4;
      if (FIX || c < D) {
        const float4 d = lds4(slot, dy_dtype, c);
        const float4 xv = lds4(slot + dy_b, x_dtype, c);
        const float4 ww = *(const float4*)(w_s + c);
        const float4 xh = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        const float4 g = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
        s1 += g.x + g.y + g.z + g.w;
        s2 += g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w;
        aw[i].x += d.x * xh.x; aw[i].y += d.y * xh.y; aw[i].z += d.z * xh.z; aw[i].w += d.w * xh.w;
        ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // the two reductions interleaved: one shuffle latency per level, not two
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= (float)D;
    s2 /= (float)D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {   // pass 2: dx (g and xhat recomputed from the staged row, bit-identical to pass 1)
      const int c = (i * 32 + lane) * 4;
      if (FIX || c < D) {
        const float4 d = lds4(slot, dy_dtype, c);
        const float4 xv = lds4(slot + dy_b, x_dtype, c);
        const float4 ww = *(const float4*)(w_s + c);
        const float4 xh = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        const float4 g = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
        float4 o;
        o.x = rstd * (g.x - s1 - xh.x * s2);
        o.y = rstd * (g.y - s1 - xh.y * s2);
        o.z = rstd * (g.z - s1 - xh.z * s2);
        o.w = rstd * (g.w - s1 - xh.w * s2);
        if (add_in) {
          const float4 a = lds4(slot + dy_b + x_b, add_dtype, c);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        st4(dx, dx_dtype, base + c, o);
        if (DENSE) {
          if (drop_p > 0.f) {   // D is even and c a multiple of 4: the four elements are two whole hash pairs
            const uint64_t pr = (uint64_t)(base + c) >> 1;
            const uint32_t k0 = drop_keep_pair(seed_mix, (uint32_t)pr, (uint32_t)(pr >> 32), thr);
            const uint32_t k1 = drop_keep_pair(seed_mix, (uint32_t)(pr + 1), (uint32_t)((pr + 1) >> 32), thr);
            o.x = (k0 & 1u) ? o.x * inv_keep : 0.f;
            o.y = (k0 & 2u) ? o.y * inv_keep : 0.f;
            o.z = (k1 & 1u) ? o.z * inv_keep : 0.f;
            o.w = (k1 & 2u) ? o.w * inv_keep : 0.f;
          }
          const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
          uint2 u;
          u.x = *(const uint32_t*)&lo;
          u.y = *(const uint32_t*)&hi;
          *(uint2*)(dx16 + base + c) = u;
          const float2 flo = __bfloat1622float2(lo), fhi = __bfloat1622float2(hi);   // the column sums see the rounded values
          ad[i].x += flo.x; ad[i].y += flo.y; ad[i].z += fhi.x; ad[i].w += fhi.y;
        }
      }
    }
    __syncwarp();                                   // every lane is done reading this slot
    issue(row + 2 * nw, k & 1);                     // refill it with the row after next
  }
  cp_async_wait<0>();
  if (!dw && !DENSE) return;
  // The rings are dead: reuse shared memory for the cross-warp reduction, one accumulator array at a time and `red_rows`
  // warps at a time (16 rows of D floats fit up to D = 3584; wider rows go in two or four groups).
  float* rbuf = (float*)ln_smem;
  auto reduce_into = [&](const float4* acc, float* out) {
    for (int g0 = 0; g0 < LNB_WARPS; g0 += red_rows) {
      __syncthreads();
      if (warp >= g0 && warp < g0 + red_rows) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (FIX || c < D) *(float4*)(rbuf + (warp - g0) * D + c) = acc[i];
        }
      }
      __syncthreads();
      for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float t = 0.f;
        for (int k = 0; k < red_rows; ++k) t += rbuf[k * D + c];
        atomicAdd(out + c, t);
      }
    }
  };
  if (dw) {
    reduce_into(aw, dw);
    reduce_into(ab, db);
  }
  if (DENSE && dbias) reduce_into(ad, dbias);
}

// -------------------------------------------------------------------------------- LayerScale backward
// Forward (GEMM epilogue): x_out = x_in + rs[row/rpg] * gamma * z, z = acc + bias (saved as bf16).
// Backward: dz = dx_out * gamma * rs (bf16 out), dgamma += sum_rows dx_out * rs * z, dbias += sum_rows dz.
// Reference: beit2.py:204-205 (gamma_1 / gamma_2, DropPath).
// A row's loads are all issued before anything consumes them (a separate, branch-free load phase): with the loads inside
// the compute loop the compiler kept one DRAM round trip per 128 columns (ncu r02j: 78 % of the stalls on the first use
// of each load, 38 % of DRAM bandwidth).
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, NV <= 8 ? 2 : 1)
layerscale_bwd_kernel(const float* __restrict__ dxo, const bf16* __restrict__ z, const float* __restrict__ gamma,
                      const float* __restrict__ rs, int rpg, bf16* __restrict__ dz, float* __restrict__ dgamma,
                      float* __restrict__ dbias, int M, int D, int rows_per_cta) {
  extern __shared__ float red[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int row0 = blockIdx.x * rows_per_cta;
  const int row1 = min(M, row0 + rows_per_cta);
  float* gs = red + 2 * LN_WARPS * D;   // gamma, zero-padded to NV * 128 columns: the compute phase needs no bounds branch
  for (int c = threadIdx.x; c < NV * 128; c += LN_WARPS * 32) gs[c] = c < D ? gamma[c] : 0.f;
  __syncthreads();
  for (int row = row0 + warp; row < row1; row += LN_WARPS) {
    const size_t base = (size_t)row * D;
    const float s = rs ? rs[row / rpg] : 1.f;
    float4 d[NV];
    uint2 zr[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const bool in = c < D;
      d[i] = in ? __ldg((const float4*)(dxo + base + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      zr[i] = in ? __ldg((const uint2*)(z + base + c)) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float2 z0 = __bfloat1622float2(*(const __nv_bfloat162*)&zr[i].x), z1 = __bfloat1622float2(*(const __nv_bfloat162*)&zr[i].y);
      const float4 gm = *(const float4*)(gs + c);
      const float4 ds = make_float4(d[i].x * s, d[i].y * s, d[i].z * s, d[i].w * s);
      const float4 o = make_float4(ds.x * gm.x, ds.y * gm.y, ds.z * gm.z, ds.w * gm.w);
      if (c < D) st4(dz, 0, base + c, o);
      ag[i].x += ds.x * z0.x; ag[i].y += ds.y * z0.y; ag[i].z += ds.z * z1.x; ag[i].w += ds.w * z1.y;
      ab[i].x += o.x; ab[i].y += o.y; ab[i].z += o.z; ab[i].w += o.w;
    }
  }
  float* rg = red;
  float* rb = red + LN_WARPS * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < D) {
      *(float4*)(rg + warp * D + c) = ag[i];
      *(float4*)(rb + warp * D + c) = ab[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < LN_WARPS; ++k) {
      sg += rg[k * D + c];
      sb += rb[k * D + c];
    }
    atomicAdd(dgamma + c, sg);
    if (dbias) atomicAdd(dbias + c, sb);
  }
}

// -------------------------------------------------------------------------------- column sum (bias grads)
// out[c] += sum_rows in[row, c]; in is bf16 [M, N] with leading dimension ld.  blockDim = (32, 8):
// each thread owns 8 consecutive columns (one 16-byte load), 8 row-lanes per CTA, rows strided over grid.y.
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ in, int64_t ld, float* __restrict__ out, int M, int N, int rows_per_cta) {
  __shared__ float red[8][256 + 8];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int row0 = blockIdx.y * rows_per_cta;
  const int row1 = min(M, row0 + rows_per_cta);
  if (c < N) {
    // four independent 16-byte loads in flight per thread (one per iteration left ~8 KB in flight per SM: 0.52 of the copy
    // bandwidth, profiles/r01_dev_kernels.jsonl)
    int r = row0 + threadIdx.y;
    for (; r + 24 < row1; r += 32) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = *(const uint4*)(in + (int64_t)(r + 8 * k) * ld + c);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162* h = (const __nv_bfloat162*)&u[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(h[q]);
          acc[2 * q] += f.x;
          acc[2 * q + 1] += f.y;
        }
      }
    }
    for (; r < row1; r += 8) {
      const uint4 u = *(const uint4*)(in + (int64_t)r * ld + c);
      const __nv_bfloat162* h = (const __nv_bfloat162*)&u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        acc[2 * q] += f.x;
        acc[2 * q + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;  // 0..255 -> one column each
  const int col = blockIdx.x * 256 + t;
  if (col < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][t];
    atomicAdd(out + col, s);
  }
}

// -------------------------------------------------------------------------------- casts / fills
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = *(const float4*)(in + 4 * i);
    st4(out, 0, 4 * i, v);
  }
}
__global__ void cast_bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    *(float4*)(out + 4 * i) = ld4(in, 0, 4 * i);
}
// Split-precision operand for an fp32-grade GEMM on the bf16 tensor cores: x = h + m + l with h = bf16(x), m = bf16(x - h),
// l = bf16(x - h - m) (24 mantissa bits in three bf16 terms).  Each f32 row [K] becomes a bf16 row [6K] of six K-blocks;
// role 0 (left operand) lays out l,m,h,m,h,h and role 1 (right operand) h,m,l,h,m,h, so one K' = 6K GEMM sums the six
// products of order <= 2 in the order lh, mm, hl, mh, hm, hh.  SMALLEST TERMS FIRST: the tensor core adds into its fp32
// accumulator with truncation, an error of ~2^-24 |acc| per k-step; with the dominant hh block last only its K/16 steps
// see the full accumulator magnitude (measured at K = 768: 1.0e-5 of max|C| with hh first, ~1e-6 with hh last = what a
// cuBLAS fp32 GEMM gives).
// Used where the reference forces fp32 inside an otherwise reduced-precision region (model_vqkd.py:154-155).
// act: 0 none, 1 tanh applied before the split (encode_task_layer's nn.Tanh, model_vqkd.py:86-90).
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ in, bf16* __restrict__ out, int M, int K, int role, int act) {
  const size_t n4 = (size_t)M * (K / 4);
  const int k4 = K / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / k4;
    const int col = (int)(i - row * k4) * 4;
    float4 v = *(const float4*)(in + row * K + col);
    if (act == 1) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
    const float x[4] = {v.x, v.y, v.z, v.w};
    bf16 t[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bf16 h = __float2bfloat16_rn(x[e]);
      const float r1 = x[e] - __bfloat162float(h);
      const bf16 m = __float2bfloat16_rn(r1);
      const bf16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
      t[0][e] = h; t[1][e] = m; t[2][e] = l;
    }
    const int sel0[6] = {2, 1, 0, 1, 0, 0}, sel1[6] = {0, 1, 2, 0, 1, 0};
    bf16* o = out + row * (size_t)(6 * K) + col;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const int w = role == 0 ? sel0[b] : sel1[b];
      uint2 u;
      __nv_bfloat162 lo = __halves2bfloat162(t[w][0], t[w][1]), hi = __halves2bfloat162(t[w][2], t[w][3]);
      u.x = *(uint32_t*)&lo; u.y = *(uint32_t*)&hi;
      *(uint2*)(o + (size_t)b * K) = u;
    }
  }
}

// out = in * (*scalar) ; in/out same dtype, out may alias in (applies the upstream loss gradient to the gradients a fused
// loss kernel prepared in its forward; out-of-place so that a retained graph can be back-propagated twice)
__global__ void scale_by_scalar_kernel(const void* in, void* out, int dtype, const float* __restrict__ scalar, size_t n4) {
  const float s = *scalar;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = ld4(in, dtype, 4 * i);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    st4(out, dtype, 4 * i, v);
  }
}

// y = gelu(x) (bf16 -> bf16) and dx = dy * gelu'(x): the two GELU sites that do not sit behind a GEMM epilogue
// (ITM head: Linear -> LayerNorm -> GELU, xfm.py:115-121; LM head backward: dense -> GELU -> LayerNorm, xroberta.py:1325-1328).
__global__ void gelu_fwd_kernel(const void* __restrict__ x, int x_dtype, bf16* __restrict__ y, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = ld4(x, x_dtype, 4 * i);
    v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
    st4(y, 0, 4 * i, v);
  }
}
__global__ void gelu_bwd_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ x, int x_dtype,
                                bf16* __restrict__ dx, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 g = ld4(dy, dy_dtype, 4 * i), v = ld4(x, x_dtype, 4 * i);
    st4(dx, 0, 4 * i, make_float4(g.x * gelu_erf_grad(v.x), g.y * gelu_erf_grad(v.y), g.z * gelu_erf_grad(v.z),
                                  g.w * gelu_erf_grad(v.w)));
  }
}
// y[i] = keep(seed, i) ? x[i] / (1-p) : 0 with the SAME (seed, row * N + col) indexing as the GEMM dropout epilogue, so
// the backward pass re-applies the forward mask to the incoming gradient without storing it (xroberta.py:302,383).
__global__ void dropout_apply_kernel(const void* __restrict__ x, int x_dtype, bf16* __restrict__ y, size_t n4, float p,
                                     uint64_t seed, const uint64_t* __restrict__ salt) {
  seed += *salt;
  const float inv_keep = 1.0f / (1.0f - p);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = ld4(x, x_dtype, 4 * i);
    v.x = drop_keep_idx(seed, 4 * i, p) ? v.x * inv_keep : 0.f;
    v.y = drop_keep_idx(seed, 4 * i + 1, p) ? v.y * inv_keep : 0.f;
    v.z = drop_keep_idx(seed, 4 * i + 2, p) ? v.z * inv_keep : 0.f;
    v.w = drop_keep_idx(seed, 4 * i + 3, p) ? v.w * inv_keep : 0.f;
    st4(y, 0, 4 * i, v);
  }
}

// -------------------------------------------------------------------------------- RoBERTa embeddings
// pos_ids = cumsum(ids != pad) * (ids != pad) + pad (xroberta.py:1747-1757); emb = word + type0 + pos; LayerNorm.
// One CTA per sequence, one warp per token (round-robin).  Saves pos_ids and LN stats for the backward.
__global__ void __launch_bounds__(LN_WARPS * 32)
roberta_embed_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
                         const float* __restrict__ type0, const float* __restrict__ w, const float* __restrict__ b,
                         bf16* __restrict__ y, float* __restrict__ pre_ln, float* __restrict__ stats,
                         int32_t* __restrict__ pos_ids, int L, int D, int pad_id, int absolute_pos, float eps) {
  extern __shared__ int s_pos[];  // [L]
  const int seq = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    int carry = 0;
    for (int t0 = 0; t0 < L; t0 += 32) {
      const int t = t0 + lane;
      const int m = (t < L && ids[(size_t)seq * L + t] != pad_id) ? 1 : 0;
      int inc = m;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      if (t < L) {
        // RoBERTa: cumsum(ids != pad) * (ids != pad) + pad (xroberta.py:1747-1757); BERT: position_ids[:, :L] (xbert.py:199-200)
        const int p = absolute_pos ? t : (carry + inc) * m + pad_id;
        s_pos[t] = p;
        pos_ids[(size_t)seq * L + t] = p;
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  for (int t = warp; t < L; t += LN_WARPS) {
    const size_t row = (size_t)seq * L + t;
    const int64_t id = ids[row];
    const float* wr = word + (size_t)id * D;
    const float* pr = pos + (size_t)s_pos[t] * D;
    float4 v[LN_MAX_VEC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float4 a = *(const float4*)(wr + c), t4 = *(const float4*)(type0 + c), p4 = *(const float4*)(pr + c);
        v[i] = make_float4((a.x + t4.x) + p4.x, (a.y + t4.y) + p4.y, (a.z + t4.z) + p4.z, (a.w + t4.w) + p4.w);
        sum += v[i].x + v[i].y + v[i].z + v[i].w;
        if (pre_ln) *(float4*)(pre_ln + row * D + c) = v[i];
      }
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        sq += a * a + bb * bb + cc * cc + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float4 ww = *(const float4*)(w + c), bb = *(const float4*)(b + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * ww.x + bb.x;
        o.y = (v[i].y - mean) * rstd * ww.y + bb.y;
        o.z = (v[i].z - mean) * rstd * ww.z + bb.z;
        o.w = (v[i].w - mean) * rstd * ww.w + bb.w;
        st4(y, 0, row * D + c, o);
      }
    }
    if (lane == 0 && stats) {
      stats[2 * row] = mean;
      stats[2 * row + 1] = rstd;
    }
  }
}

// Scatter-add the (pre-LayerNorm) embedding gradient into the word / position / token-type tables.
// Padding rows receive gradients exactly like torch's nn.Embedding without padding_idx masking would NOT:
// nn.Embedding(padding_idx=pad) zeroes the gradient of row `pad` (xroberta.py:80,100-102), reproduced here.
__global__ void __launch_bounds__(256)
roberta_embed_bwd_kernel(const float* __restrict__ dpre, const int64_t* __restrict__ ids,
                         const int32_t* __restrict__ pos_ids, float* __restrict__ dword, float* __restrict__ dpos,
                         float* __restrict__ dtype0, int rows, int D, int word_pad, int pos_pad) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const int64_t id = ids[row];
  const int p = pos_ids[row];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float g = dpre[(size_t)row * D + c];
    if (id != word_pad) atomicAdd(dword + (size_t)id * D + c, g);   // nn.Embedding(padding_idx) rows get no gradient
    if (p != pos_pad) atomicAdd(dpos + (size_t)p * D + c, g);
    atomicAdd(dtype0 + c, g);
  }
}

// -------------------------------------------------------------------------------- ViT patch pipeline
// im2col for the 16x16 stride-16 conv (beit2.py:229): out[b*np + p, c*P*P + py*P + px] = image[b, c, gy*P+py, gx*P+px].
// With pre_mul_ptr (a device scalar m) the VQ-KD pre-processing x * m / 127.5 - 1 (model_vqkd.py:125-131) is applied.
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int C, int H, int W, int P,
              const float* __restrict__ pre_mul_ptr) {
  const float pre_mul = pre_mul_ptr ? __ldg(pre_mul_ptr) : 0.f;
  const int gw = W / P, gh = H / P;
  const int K = C * P * P;
  const size_t total4 = (size_t)B * gh * gw * K / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 4;
    const int k = (int)(e % K);
    const size_t rowp = e / K;
    const int p = (int)(rowp % (gh * gw));
    const int b = (int)(rowp / (gh * gw));
    const int c = k / (P * P), py = (k / P) % P, px = k % P;  // px multiple of 4
    const int gy = p / gw, gx = p % gw;
    float4 v = *(const float4*)(img + (((size_t)b * C + c) * H + gy * P + py) * W + gx * P + px);
    if (pre_mul != 0.f) {
      v.x = v.x * pre_mul / 127.5f - 1.0f; v.y = v.y * pre_mul / 127.5f - 1.0f;
      v.z = v.z * pre_mul / 127.5f - 1.0f; v.w = v.w * pre_mul / 127.5f - 1.0f;
    }
    st4(out, 0, e, v);
  }
}

// x[b, 0] = cls (+ pos[0]); x[b, 1+p] = (mask[b,p] ? mask_token : patch[b,p]) (+ pos[1+p]).  (beit2.py:438-449,
// vqkd_vit.py:378-385).  patch is the f32 conv output [B*np, D]; x is the f32 residual stream [B*(np+1), D].
__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ mask_token,
                       const uint8_t* __restrict__ mask, const float* __restrict__ pos, float* __restrict__ x, int B,
                       int np, int D) {
  const int row = blockIdx.x;  // b * (np + 1) + t
  const int b = row / (np + 1), t = row % (np + 1);
  const float* src;
  if (t == 0) src = cls;
  else if (mask && mask[(size_t)b * np + t - 1]) src = mask_token;
  else src = patch + ((size_t)b * np + t - 1) * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 v = *(const float4*)(src + c);
    if (pos) {
      const float4 p4 = *(const float4*)(pos + (size_t)t * D + c);
      v.x += p4.x; v.y += p4.y; v.z += p4.z; v.w += p4.w;
    }
    *(float4*)(x + (size_t)row * D + c) = v;
  }
}

// Backward of assemble: dpatch[b,p] = mask ? 0 : dx[b,1+p] (bf16, feeds the conv wgrad GEMM);
// dcls += sum_b dx[b,0]; dmask_token += sum over masked positions.
__global__ void __launch_bounds__(256)
assemble_tokens_bwd_kernel(const float* __restrict__ dx, const uint8_t* __restrict__ mask, bf16* __restrict__ dpatch,
                           float* __restrict__ dcls, float* __restrict__ dmask_token, int B, int np, int D) {
  // grid (B, 4): a CTA walks every 4th token row of one image, thread = float4 column; the mask-token gradient is summed in
  // registers and reaches global memory as ONE atomic per column and CTA (the first version issued one per masked element:
  // 5.5 M atomics on 768 addresses at B = 96, 152 us).
  const int b = blockIdx.x;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 am = make_float4(0.f, 0.f, 0.f, 0.f);
    bool any = false;
    for (int t = blockIdx.y; t <= np; t += gridDim.y) {
      const size_t row = (size_t)b * (np + 1) + t;
      const float4 v = *(const float4*)(dx + row * D + c);
      if (t == 0) {
        atomicAdd(dcls + c, v.x); atomicAdd(dcls + c + 1, v.y); atomicAdd(dcls + c + 2, v.z); atomicAdd(dcls + c + 3, v.w);
      } else {
        const bool masked = mask && mask[(size_t)b * np + t - 1];
        if (masked) {
          am.x += v.x; am.y += v.y; am.z += v.z; am.w += v.w;
          any = true;
        }
        st4(dpatch, 0, ((size_t)b * np + t - 1) * D + c, masked ? make_float4(0.f, 0.f, 0.f, 0.f) : v);
      }
    }
    if (any) {
      atomicAdd(dmask_token + c, am.x); atomicAdd(dmask_token + c + 1, am.y);
      atomicAdd(dmask_token + c + 2, am.z); atomicAdd(dmask_token + c + 3, am.w);
    }
  }
}

// Mean-pool pseudo-CLS (beit2.py:456-466): y[b,0,:] = mean_p y[b,1+p,:], written to the bf16 and f32 copies.
// grid (B, D / 128): a CTA owns 128 columns of one image; thread = (row group 0..7, float4 column), 8 partial sums per column
// reduced through shared memory in a fixed order (the first version walked the 196 rows serially per thread: 224 us at
// B = 96, two launches per step).
__global__ void __launch_bounds__(256)
meanpool_fwd_kernel(bf16* __restrict__ y, float* __restrict__ y32, int np, int D) {
  __shared__ float4 part[8][32];
  const int b = blockIdx.x;
  const int c = blockIdx.y * 128 + (threadIdx.x & 31) * 4;
  const int rg = threadIdx.x >> 5;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < D) {
    const float* base = y32 + (size_t)b * (np + 1) * D + c;
    for (int p = 1 + rg; p <= np; p += 8) {
      const float4 v = *(const float4*)(base + (size_t)p * D);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  part[rg][threadIdx.x & 31] = s;
  __syncthreads();
  if (rg == 0 && c < D) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = part[k][threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float inv = 1.0f / (float)np;
    s = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
    *(float4*)(y32 + (size_t)b * (np + 1) * D + c) = s;
    st4(y, 0, (size_t)b * (np + 1) * D + c, s);
  }
}
// dy_ln[b,t,:] = t == 0 ? 0 : dout[b,t,:] + dout[b,0,:] / np
__global__ void __launch_bounds__(256)
meanpool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dy, int np, int D) {
  const int row = blockIdx.x;
  const int b = row / (np + 1), t = row % (np + 1);
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t > 0) {
      const float4 a = *(const float4*)(dout + (size_t)row * D + c);
      const float4 m = *(const float4*)(dout + (size_t)b * (np + 1) * D + c);
      const float inv = 1.0f / (float)np;
      o = make_float4(a.x + m.x * inv, a.y + m.y * inv, a.z + m.z * inv, a.w + m.w * inv);
    }
    *(float4*)(dy + (size_t)row * D + c) = o;
  }
}

// -------------------------------------------------------------------------------- row gather / scatter-add
// out[i, :] = in[index[i], :]   (dtype generic; D multiple of 4)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const void* __restrict__ in, int in_dtype, const int64_t* __restrict__ index, void* __restrict__ out,
                   int out_dtype, int D) {
  const size_t src = (size_t)index[blockIdx.x] * D, dst = (size_t)blockIdx.x * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) st4(out, out_dtype, dst + c, ld4(in, in_dtype, src + c));
}
// out[index[i], :] += in[i, :]   (out f32, atomics: several i may share a destination)
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const void* __restrict__ in, int in_dtype, const int64_t* __restrict__ index,
                        float* __restrict__ out, int D) {
  const size_t dst = (size_t)index[blockIdx.x] * D, src = (size_t)blockIdx.x * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    const float4 v = ld4(in, in_dtype, src + c);
    atomicAdd(out + dst + c, v.x); atomicAdd(out + dst + c + 1, v.y);
    atomicAdd(out + dst + c + 2, v.z); atomicAdd(out + dst + c + 3, v.w);
  }
}

// Relative-position bias: bias[h, i, j] = table[index[i, j], h] (beit2.py:139-144), and its scatter-add backward.
// bias / dbias rows have leading dimension ld (>= N) so the attention kernels can use aligned rows.
__global__ void __launch_bounds__(256)
relpos_bias_fwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ index, float* __restrict__ bias,
                       int N, int ld, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * N) return;
  const int64_t r = index[i];
  const size_t o = (size_t)(i / N) * ld + (i % N);
  for (int h = 0; h < H; ++h) bias[(size_t)h * N * ld + o] = table[r * H + h];
}
__global__ void __launch_bounds__(256)
relpos_bias_bwd_kernel(const float* __restrict__ dbias, const int64_t* __restrict__ index, float* __restrict__ dtable,
                       int N, int ld, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * N) return;
  const int64_t r = index[i];
  const size_t o = (size_t)(i / N) * ld + (i % N);
  for (int h = 0; h < H; ++h) atomicAdd(dtable + r * H + h, dbias[(size_t)h * N * ld + o]);
}

// Sum over the batch of the bf16 dS dump from attention backward: out[h,i,j] = sum_b ds[b,h,i,j].
__global__ void __launch_bounds__(256)
batch_sum_bf16_kernel(const bf16* __restrict__ in, float* __restrict__ out, int B, size_t per) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= per) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 v = ld4(in, 0, (size_t)b * per + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *(float4*)(out + i) = s;
}

// ------------------------------------------------------------------------------------------ host
// LN_DISPATCH covers rows up to 4096 wide (NV = 32: the 3072-wide hidden layer of build_mlp on concatenated CLS rows,
// model_nlvr.py:25); the embedding kernel keeps its row in LN_MAX_VEC registers (<= 2048).
static int ln_check(int D, int max_vec = 32) {
  if (D % 4 != 0 || D > 32 * 4 * max_vec) {
    set_error("feature dim %d unsupported (need multiple of 4, <= %d)", D, 32 * 4 * max_vec);
    return XFM_ERR_BAD_ARG;
  }
  return 0;
}
static int grid_1d(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
// Per-lane float4 slots actually needed for D features: kernels are instantiated per slot count so the register
// arrays match D (a single 16-slot instantiation needed 255 registers and spilled in the backward kernel).
#define LN_DISPATCH(D, ...)                          \
  do {                                               \
    const int nv_ = ((D) + 127) / 128;               \
    if (nv_ <= 2) { constexpr int NV = 2; __VA_ARGS__; }        \
    else if (nv_ <= 4) { constexpr int NV = 4; __VA_ARGS__; }   \
    else if (nv_ <= 6) { constexpr int NV = 6; __VA_ARGS__; }   \
    else if (nv_ <= 8) { constexpr int NV = 8; __VA_ARGS__; }   \
    else if (nv_ <= 12) { constexpr int NV = 12; __VA_ARGS__; } \
    else if (nv_ <= 16) { constexpr int NV = 16; __VA_ARGS__; } \
    else if (nv_ <= 24) { constexpr int NV = 24; __VA_ARGS__; } \
    else { constexpr int NV = 32; __VA_ARGS__; }                \
  } while (0)

#define LAUNCH_END()  \
  count_launch();     \
  return (int)cudaGetLastError();

int layernorm_fwd(const void* x, int x_dtype, const float* w, const float* b, void* y, int y_dtype, float* y2, float* stats,
                  int M, int D, float eps, cudaStream_t s) {
  if (ln_check(D)) return XFM_ERR_BAD_ARG;
  if (M <= 0) return 0;
  LN_DISPATCH(D, (layernorm_fwd_kernel<NV><<<(M + LN_WARPS - 1) / LN_WARPS, LN_WARPS * 32, 0, s>>>(x, x_dtype, w, b, y, y_dtype, y2,
                                                                                               stats, M, D, eps)));
  LAUNCH_END();
}

int layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* stats, const float* w,
                  const void* add_in, int add_dtype, void* dx, int dx_dtype, float* dw, float* db, int M, int D,
                  cudaStream_t s) {
  return layernorm_bwd_dense(dy, dy_dtype, x, x_dtype, stats, w, add_in, add_dtype, dx, dx_dtype, dw, db, nullptr, nullptr, 0.f, 0,
                             M, D, s);
}

int layernorm_bwd_dense(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* stats, const float* w,
                        const void* add_in, int add_dtype, void* dx, int dx_dtype, float* dw, float* db, bf16* dx16,
                        float* dbias, float drop_p, uint64_t drop_seed, int M, int D, cudaStream_t s) {
  if (ln_check(D)) return XFM_ERR_BAD_ARG;
  if (dx16 && (((uintptr_t)dx16 & 7) || (D & 3))) {
    set_error("layernorm_bwd_dense: dx16 must be 8-byte aligned and D a multiple of 4");
    return XFM_ERR_BAD_ARG;
  }
  if (D & 7) {
    set_error("layernorm_bwd: feature dim %d must be a multiple of 8 (16-byte cp.async rows)", D);
    return XFM_ERR_BAD_ARG;
  }
  if (M <= 0) return 0;
  // one CTA of 16 warps per SM (the staging rings take ~100-220 KB), >= 2 rows per warp so the ring actually pipelines
  const size_t slot = (size_t)D * ((dy_dtype == 1 ? 4 : 2) + (x_dtype == 1 ? 4 : 2) + (add_in ? (add_dtype == 1 ? 4 : 2) : 0)) + 16;
  const size_t wbytes = (size_t)D * 4, budget = 224 * 1024;
  int nw = LNB_WARPS;   // wide fp32 rows: fewer warps get a staging ring so the rings fit shared memory
  while (nw > 1 && wbytes + (size_t)nw * 2 * slot > budget) --nw;
  int ctas = num_sms();
  int rpc = (M + ctas - 1) / ctas;
  if (rpc < 2 * nw) rpc = 2 * nw;
  const int grid = (M + rpc - 1) / rpc;
  size_t smem = wbytes + (size_t)nw * 2 * slot;
  int red_rows = LNB_WARPS;
  while (red_rows > 1 && (size_t)red_rows * D * sizeof(float) > budget) red_rows >>= 1;
  const size_t red = (size_t)red_rows * D * sizeof(float);
  if (smem < red) smem = red;
  if (smem > 227 * 1024) {
    set_error("layernorm_bwd: D=%d needs %zu bytes of shared memory", D, smem);
    return XFM_ERR_BAD_ARG;
  }
#define LNB_LAUNCH(NVv, DENSEv, DTv)                                                                                      \
  do {                                                                                                                    \
    auto kern = layernorm_bwd_kernel<NVv, DENSEv, DTv>;                                                                   \
    static size_t attr = 0; /* one per instantiation */                                                                   \
    if (smem > attr) {                                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
      if (e != cudaSuccess) return (int)e;                                                                                \
      attr = smem;                                                                                                        \
    }                                                                                                                     \
    kern<<<grid, LNB_WARPS * 32, smem, s>>>(dy, dy_dtype, x, x_dtype, stats, w, add_in, add_dtype, dx, dx_dtype, dw, db, M, D, \
                                            rpc, nw, red_rows, dx16, dbias, drop_p, drop_seed, seed_salt_ptr());          \
  } while (0)
  if (!dx16) {
    dbias = nullptr;
    drop_p = 0.f;
    drop_seed = 0;
  }
  if (D == 768 && x_dtype == 1 && dx_dtype == 1) {   // the shapes of the base model's steps: dtypes baked in
    const int dt = (dy_dtype == 1 ? 1 : 0) | 2 | ((add_in ? (add_dtype == 1 ? 2 : 1) : 0) << 2) | 16;
    bool done = true;
    if (dx16) {
      switch (dt) {
        case 18: LNB_LAUNCH(6, true, 18); break;    // dy bf16
        case 19: LNB_LAUNCH(6, true, 19); break;    // dy fp32
        default: done = false;
      }
    } else {
      switch (dt) {
        case 18: LNB_LAUNCH(6, false, 18); break;   // dy bf16, no add_in
        case 19: LNB_LAUNCH(6, false, 19); break;   // dy fp32, no add_in
        case 26: LNB_LAUNCH(6, false, 26); break;   // dy bf16, add_in fp32 (BEiT blocks)
        case 27: LNB_LAUNCH(6, false, 27); break;   // dy fp32, add_in fp32
        default: done = false;
      }
    }
    if (done) {
      LAUNCH_END();
    }
  }
  if (dx16) {
    LN_DISPATCH(D, LNB_LAUNCH(NV, true, -1));
    LAUNCH_END();
  }
  LN_DISPATCH(D, LNB_LAUNCH(NV, false, -1));
  LAUNCH_END();
#undef LNB_LAUNCH
}

int layerscale_bwd(const float* dxo, const bf16* z, const float* gamma, const float* rs, int rpg, bf16* dz, float* dgamma,
                   float* dbias, int M, int D, cudaStream_t s) {
  if (ln_check(D)) return XFM_ERR_BAD_ARG;
  if (M <= 0) return 0;
  const int ctas = num_sms() * (D <= 1024 ? 2 : 1);   // the resident CTAs (register-bound): exactly one wave
  int rpc = (M + ctas - 1) / ctas;
  if (rpc < LN_WARPS) rpc = LN_WARPS;
  const int grid = (M + rpc - 1) / rpc;
  LN_DISPATCH(D, {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(layerscale_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * LN_WARPS + 1) * NV * 128 * 4);
      attr = true;
    }
    layerscale_bwd_kernel<NV><<<grid, LN_WARPS * 32, ((size_t)2 * LN_WARPS * D + NV * 128) * sizeof(float), s>>>(
        dxo, z, gamma, rs, rpg > 0 ? rpg : 1, dz, dgamma, dbias, M, D, rpc);
  });
  LAUNCH_END();
}

int colsum_bf16(const bf16* in, int64_t ld, float* out, int M, int N, cudaStream_t s) {
  if ((N & 7) || (ld & 7)) {
    set_error("colsum: N and ld must be multiples of 8");
    return XFM_ERR_BAD_ARG;
  }
  if (M <= 0) return 0;
  const int gx = (N + 255) / 256;
  int gy = (num_sms() * 4 + gx - 1) / gx;
  int rpc = (M + gy - 1) / gy;
  rpc = ((rpc + 7) / 8) * 8;
  gy = (M + rpc - 1) / rpc;
  colsum_kernel<<<dim3(gx, gy), dim3(32, 8), 0, s>>>(in, ld, out, M, N, rpc);
  LAUNCH_END();
}

int cast_f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t s) {
  if (n & 3) { set_error("cast: n must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  if (!n) return 0;
  cast_f32_to_bf16_kernel<<<grid_1d(n / 4, 256), 256, 0, s>>>(in, out, n / 4);
  LAUNCH_END();
}
int cast_bf16_to_f32(const bf16* in, float* out, size_t n, cudaStream_t s) {
  if (n & 3) { set_error("cast: n must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  if (!n) return 0;
  cast_bf16_to_f32_kernel<<<grid_1d(n / 4, 256), 256, 0, s>>>(in, out, n / 4);
  LAUNCH_END();
}
int split_bf16x3(const float* in, bf16* out, int M, int K, int role, int act, cudaStream_t s) {
  if (K & 3 || M < 0 || (role != 0 && role != 1)) { set_error("split_bf16x3: K must be a multiple of 4, role 0/1"); return XFM_ERR_BAD_ARG; }
  if (!M) return 0;
  split_bf16x3_kernel<<<grid_1d((size_t)M * (K / 4), 256), 256, 0, s>>>(in, out, M, K, role, act);
  LAUNCH_END();
}
int scale_by_scalar(const void* in, void* out, int dtype, const float* scalar, size_t n, cudaStream_t s) {
  if (n & 3) { set_error("scale: n must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  if (!n) return 0;
  scale_by_scalar_kernel<<<grid_1d(n / 4, 256), 256, 0, s>>>(in, out, dtype, scalar, n / 4);
  LAUNCH_END();
}

int gelu_fwd(const void* x, int x_dtype, bf16* y, size_t n, cudaStream_t s) {
  if (n & 3) { set_error("gelu: n must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  if (!n) return 0;
  gelu_fwd_kernel<<<grid_1d(n / 4, 256), 256, 0, s>>>(x, x_dtype, y, n / 4);
  LAUNCH_END();
}
int gelu_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, bf16* dx, size_t n, cudaStream_t s) {
  if (n & 3) { set_error("gelu: n must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  if (!n) return 0;
  gelu_bwd_kernel<<<grid_1d(n / 4, 256), 256, 0, s>>>(dy, dy_dtype, x, x_dtype, dx, n / 4);
  LAUNCH_END();
}
int dropout_apply(const void* x, int x_dtype, bf16* y, size_t n, float p, uint64_t seed, cudaStream_t s) {
  if (n & 3) { set_error("dropout: n must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  if (!n) return 0;
  dropout_apply_kernel<<<grid_1d(n / 4, 256), 256, 0, s>>>(x, x_dtype, y, n / 4, p, seed, seed_salt_ptr());
  LAUNCH_END();
}

int roberta_embed_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, const float* w,
                      const float* b, bf16* y, float* pre_ln, float* stats, int32_t* pos_ids, int B, int L, int D,
                      int pad_id, int absolute_pos, float eps, cudaStream_t s) {
  if (ln_check(D, LN_MAX_VEC)) return XFM_ERR_BAD_ARG;
  if (B <= 0) return 0;
  roberta_embed_fwd_kernel<<<B, LN_WARPS * 32, L * sizeof(int), s>>>(ids, word, pos, type0, w, b, y, pre_ln, stats, pos_ids, L, D,
                                                                   pad_id, absolute_pos, eps);
  LAUNCH_END();
}
int roberta_embed_bwd(const float* dpre, const int64_t* ids, const int32_t* pos_ids, float* dword, float* dpos,
                      float* dtype0, int rows, int D, int word_pad, int pos_pad, cudaStream_t s) {
  if (rows <= 0) return 0;
  roberta_embed_bwd_kernel<<<rows, 256, 0, s>>>(dpre, ids, pos_ids, dword, dpos, dtype0, rows, D, word_pad, pos_pad);
  LAUNCH_END();
}

int im2col(const float* img, bf16* out, int B, int C, int H, int W, int P, const float* pre_mul, cudaStream_t s) {
  if (P % 4 || H % P || W % P) { set_error("im2col: bad geometry"); return XFM_ERR_BAD_ARG; }
  const size_t n4 = (size_t)B * C * H * W / 4;
  im2col_kernel<<<grid_1d(n4, 256), 256, 0, s>>>(img, out, B, C, H, W, P, pre_mul);
  LAUNCH_END();
}
int assemble_tokens(const float* patch, const float* cls, const float* mask_token, const uint8_t* mask, const float* pos,
                    float* x, int B, int np, int D, cudaStream_t s) {
  assemble_tokens_kernel<<<B * (np + 1), 256, 0, s>>>(patch, cls, mask_token, mask, pos, x, B, np, D);
  LAUNCH_END();
}
int assemble_tokens_bwd(const float* dx, const uint8_t* mask, bf16* dpatch, float* dcls, float* dmask_token, int B, int np,
                        int D, cudaStream_t s) {
  assemble_tokens_bwd_kernel<<<dim3(B, 4), 256, 0, s>>>(dx, mask, dpatch, dcls, dmask_token, B, np, D);
  LAUNCH_END();
}
int meanpool_fwd(bf16* y, float* y32, int B, int np, int D, cudaStream_t s) {
  if (D & 3) { set_error("meanpool: D must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  meanpool_fwd_kernel<<<dim3(B, (D + 127) / 128), 256, 0, s>>>(y, y32, np, D);
  LAUNCH_END();
}
int meanpool_bwd(const float* dout, float* dy, int B, int np, int D, cudaStream_t s) {
  meanpool_bwd_kernel<<<B * (np + 1), 256, 0, s>>>(dout, dy, np, D);
  LAUNCH_END();
}
int gather_rows(const void* in, int in_dtype, const int64_t* index, void* out, int out_dtype, int n, int D, cudaStream_t s) {
  if (D & 3) { set_error("gather: D must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  if (n <= 0) return 0;
  gather_rows_kernel<<<n, 256, 0, s>>>(in, in_dtype, index, out, out_dtype, D);
  LAUNCH_END();
}
int scatter_add_rows(const void* in, int in_dtype, const int64_t* index, float* out, int n, int D, cudaStream_t s) {
  if (D & 3) { set_error("scatter: D must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  if (n <= 0) return 0;
  scatter_add_rows_kernel<<<n, 256, 0, s>>>(in, in_dtype, index, out, D);
  LAUNCH_END();
}
int relpos_bias_fwd(const float* table, const int64_t* index, float* bias, int N, int ld, int H, cudaStream_t s) {
  relpos_bias_fwd_kernel<<<(N * N + 255) / 256, 256, 0, s>>>(table, index, bias, N, ld, H);
  LAUNCH_END();
}
int relpos_bias_bwd(const float* dbias, const int64_t* index, float* dtable, int N, int ld, int H, cudaStream_t s) {
  relpos_bias_bwd_kernel<<<(N * N + 255) / 256, 256, 0, s>>>(dbias, index, dtable, N, ld, H);
  LAUNCH_END();
}
int batch_sum_bf16(const bf16* in, float* out, int B, size_t per, cudaStream_t s) {
  if (per & 3) { set_error("batch_sum: per must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  batch_sum_bf16_kernel<<<(unsigned)((per / 4 + 255) / 256), 256, 0, s>>>(in, out, B, per);
  LAUNCH_END();
}

}  // namespace xfm
