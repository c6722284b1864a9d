// Loss-side kernels of the XFM hot path: softmax cross-entropy over wide vocabularies (MLM / MIM heads),
// the ITC contrastive loss fused with its backward, ITM hard-negative weights + on-device sampling, and the
// VQ-KD codebook argmin.
#include "common.cuh"
#include "internal.h"

namespace xfm {

// ---------------------------------------------------------------------------------------- block reductions
template <int THREADS>
XFM_DEVINL float block_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = (threadIdx.x < THREADS / 32) ? sh[threadIdx.x] : -INFINITY;
    r = warp_max(r);
    if (threadIdx.x == 0) sh[0] = r;
  }
  __syncthreads();
  const float out = sh[0];
  __syncthreads();
  return out;
}
template <int THREADS>
XFM_DEVINL float block_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = (threadIdx.x < THREADS / 32) ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) sh[0] = r;
  }
  __syncthreads();
  const float out = sh[0];
  __syncthreads();
  return out;
}

// ======================================================================================== cross-entropy
// F.cross_entropy(logits, labels, ignore_index=-100) (xroberta.py:1298-1299, xfm.py:629,795-802).
// ce_fwd: per-row lse and loss; ce_reduce: mean over non-ignored rows; ce_bwd: dlogits (bf16) scaled by the
// upstream gradient read from device memory (no host sync).
constexpr int CE_THREADS = 256;

__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int V,
              float* __restrict__ row_loss, float* __restrict__ lse) {
  __shared__ float sh[32];
  const int row = blockIdx.x;
  const int64_t label = labels[row];
  if (label < 0) {  // ignore_index
    if (threadIdx.x == 0) { row_loss[row] = 0.f; lse[row] = 0.f; }
    return;
  }
  const float* x = logits + (int64_t)row * ld;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < V; j += CE_THREADS) mx = fmaxf(mx, x[j]);
  mx = block_max<CE_THREADS>(mx, sh);
  float sum = 0.f;
  for (int j = threadIdx.x; j < V; j += CE_THREADS) sum += __expf(x[j] - mx);
  sum = block_sum<CE_THREADS>(sum, sh);
  if (threadIdx.x == 0) {
    const float l = mx + logf(sum);
    lse[row] = l;
    row_loss[row] = l - x[label];
  }
}

__global__ void __launch_bounds__(CE_THREADS)
ce_reduce_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ labels, int R, float* __restrict__ loss,
                 float* __restrict__ count) {
  __shared__ float sh[32];
  float s = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < R; i += CE_THREADS) {
    if (labels[i] >= 0) { s += row_loss[i]; c += 1.f; }
  }
  s = block_sum<CE_THREADS>(s, sh);
  c = block_sum<CE_THREADS>(c, sh);
  if (threadIdx.x == 0) {
    *loss = s / c;  // all-ignored -> nan, like torch
    *count = c;
  }
}

__global__ void __launch_bounds__(CE_THREADS)
ce_bwd_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, const float* __restrict__ lse,
              const float* __restrict__ count, const float* __restrict__ row_scale, const float* __restrict__ upstream,
              bf16* __restrict__ dlogits, int64_t ldd, int V) {
  const int row = blockIdx.x;
  const int64_t label = labels[row];
  bf16* d = dlogits + (int64_t)row * ldd;
  if (label < 0) {
    for (int j = threadIdx.x; j < (int)ldd; j += CE_THREADS) d[j] = __float2bfloat16(0.f);
    return;
  }
  // mean reduction: 1 / count; reduction='none' followed by a weighted sum (model_generation.py:130-131): row_scale[row]
  const float g = (upstream ? *upstream : 1.f) * (row_scale ? row_scale[row] : 1.f / *count);
  const float l = lse[row];
  const float* x = logits + (int64_t)row * ld;
  for (int j = threadIdx.x; j < (int)ldd; j += CE_THREADS) {
    float v = 0.f;
    if (j < V) v = (__expf(x[j] - l) - (j == label ? 1.f : 0.f)) * g;
    d[j] = __float2bfloat16(v);
  }
}

// ======================================================================================== small fp32 GEMM
// C[M,N] = alpha * sum_k A(m,k) * B(n,k) with arbitrary element strides (so any transpose combination).
// Exact fp32 FMAs; used for the ITC / hard-negative similarity matrices (<= 768^2 x 256) and their gradients.
__global__ void __launch_bounds__(256)
sgemm_small_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbn,
                   int64_t sbk, float* __restrict__ C, int64_t ldc, int M, int N, int K, const float* __restrict__ alpha_ptr,
                   float alpha_mul, int alpha_div, const float* __restrict__ bias, int accumulate) {
  __shared__ float sA[32][33], sB[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int rr = ty + 8 * r;
      const int m = m0 + rr, n = n0 + rr, k = k0 + tx;
      sA[rr][tx] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
      sB[rr][tx] = (n < N && k < K) ? B[n * sbn + k * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float bv = sB[tx][k];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sA[ty + 8 * r][k], bv, acc[r]);
    }
    __syncthreads();
  }
  float alpha = alpha_mul;
  if (alpha_ptr) alpha = alpha_div ? alpha_mul / *alpha_ptr : alpha_mul * *alpha_ptr;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + ty + 8 * r, n = n0 + tx;
    if (m < M && n < N) {
      float v = acc[r] * alpha + (bias ? bias[n] : 0.f);
      if (accumulate) v += C[(int64_t)m * ldc + n];
      C[(int64_t)m * ldc + n] = v;
    }
  }
}

static void sgemm_small(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
                        int64_t ldc, int M, int N, int K, const float* alpha_ptr, float alpha_mul, int alpha_div,
                        cudaStream_t s, const float* bias = nullptr, int accumulate = 0) {
  sgemm_small_kernel<<<dim3((N + 31) / 32, (M + 31) / 32), 256, 0, s>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, alpha_ptr,
                                                                       alpha_mul, alpha_div, bias, accumulate);
  count_launch();
}
// Exact-fp32 small GEMM for latency-bound heads (ITC projections, xfm.py:614-621): C (+)= A.B^T + bias.
int sgemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C, int64_t ldc,
              int M, int N, int K, const float* bias, int accumulate, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  sgemm_small(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, nullptr, 1.f, 0, s, bias, accumulate);
  return (int)cudaGetLastError();
}

// ======================================================================================== ITC (xfm.py:683-715)
// S = I_all T_all^T / temp (n x n).  Blocks [0,n): row i -> i2t term; blocks [n,2n): column j -> t2i term.
// Labels: arange (idx == null) or soft labels pos/pos.sum(1) with pos = (idx_i == idx_j).
constexpr int ITC_THREADS = 256;
__global__ void __launch_bounds__(ITC_THREADS)
itc_lse_loss_kernel(const float* __restrict__ S, int n, const int64_t* __restrict__ idx, float* __restrict__ row_lse,
                    float* __restrict__ col_lse, float* __restrict__ possum, float* __restrict__ loss) {
  __shared__ float sh[32];
  const bool is_col = blockIdx.x >= n;
  const int r = is_col ? blockIdx.x - n : blockIdx.x;
  const int64_t stride = is_col ? n : 1;
  const float* x = S + (is_col ? r : (int64_t)r * n);
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += ITC_THREADS) mx = fmaxf(mx, x[j * stride]);
  mx = block_max<ITC_THREADS>(mx, sh);
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += ITC_THREADS) sum += __expf(x[j * stride] - mx);
  sum = block_sum<ITC_THREADS>(sum, sh);
  const float l = mx + logf(sum);
  float term;
  if (!idx) {
    term = l - x[r * stride];
  } else {
    const int64_t me = idx[r];
    float ps = 0.f, dot = 0.f;
    for (int j = threadIdx.x; j < n; j += ITC_THREADS) {
      if (idx[j] == me) { ps += 1.f; dot += x[j * stride] - l; }
    }
    ps = block_sum<ITC_THREADS>(ps, sh);
    dot = block_sum<ITC_THREADS>(dot, sh);
    term = -dot / ps;
    if (threadIdx.x == 0 && !is_col) possum[r] = ps;  // pos is symmetric: same sum serves both directions
  }
  if (threadIdx.x == 0) {
    (is_col ? col_lse : row_lse)[r] = l;
    atomicAdd(loss, term / (2.f * (float)n));
  }
}

// dS[i,j] = (softmax_row_i[j] - lab[i,j] + softmax_col_j[i] - lab[j,i]) / (2n);  dtemp += -sum dS * S / temp.
__global__ void __launch_bounds__(ITC_THREADS)
itc_dlogits_kernel(const float* __restrict__ S, int n, const int64_t* __restrict__ idx, const float* __restrict__ row_lse,
                   const float* __restrict__ col_lse, const float* __restrict__ possum, const float* __restrict__ temp,
                   float* __restrict__ dS, float* __restrict__ dtemp) {
  __shared__ float sh[32];
  const int i = blockIdx.x;
  const float rl = row_lse[i];
  const float inv2n = 1.f / (2.f * (float)n);
  float acc = 0.f;
  for (int j = threadIdx.x; j < n; j += ITC_THREADS) {
    const float s = S[(int64_t)i * n + j];
    float lab_r, lab_c;
    if (!idx) {
      lab_r = lab_c = (i == j) ? 1.f : 0.f;
    } else {
      const float pos = (idx[i] == idx[j]) ? 1.f : 0.f;
      lab_r = pos / possum[i];
      lab_c = pos / possum[j];
    }
    const float d = ((__expf(s - rl) - lab_r) + (__expf(s - col_lse[j]) - lab_c)) * inv2n;
    dS[(int64_t)i * n + j] = d;
    acc += d * s;
  }
  acc = block_sum<ITC_THREADS>(acc, sh);
  if (threadIdx.x == 0) atomicAdd(dtemp, -acc / *temp);
}

int itc_loss_fused(const float* image_all, const float* text_all, int n, int E, const int64_t* idx_all, const float* temp,
                   int local_off, int local_n, float* work, float* loss, float* d_image_local, float* d_text_local,
                   float* dtemp, cudaStream_t s) {
  // work: S[n*n] | dS[n*n] | row_lse[n] | col_lse[n] | possum[n]
  float* S = work;
  float* dS = S + (size_t)n * n;
  float* row_lse = dS + (size_t)n * n;
  float* col_lse = row_lse + n;
  float* possum = col_lse + n;
  cudaMemsetAsync(loss, 0, sizeof(float), s);
  cudaMemsetAsync(dtemp, 0, sizeof(float), s);
  sgemm_small(image_all, E, 1, text_all, E, 1, S, n, n, n, E, temp, 1.f, 1, s);
  itc_lse_loss_kernel<<<2 * n, ITC_THREADS, 0, s>>>(S, n, idx_all, row_lse, col_lse, possum, loss);
  count_launch();
  itc_dlogits_kernel<<<n, ITC_THREADS, 0, s>>>(S, n, idx_all, row_lse, col_lse, possum, temp, dS, dtemp);
  count_launch();
  // AllGather.backward keeps only the local slice (xfm.py:93-98):
  //   dI_local[i,:] = sum_j dS[off+i, j] T[j,:] / temp ;  dT_local[j,:] = sum_i dS[i, off+j] I[i,:] / temp
  sgemm_small(dS + (size_t)local_off * n, n, 1, text_all, 1, E, d_image_local, E, local_n, E, n, temp, 1.f, 1, s);
  sgemm_small(dS + local_off, 1, n, image_all, 1, E, d_text_local, E, local_n, E, n, temp, 1.f, 1, s);
  return (int)cudaGetLastError();
}

// ======================================================================================== ITM hard negatives
// weights = softmax(sim / temp, dim=1) + 1e-5, zero where idx matches (or the diagonal); one multinomial draw per
// row (xfm.py:717-746), sampled on the device by inverse CDF so the 2B host round trips disappear.
// Blocks [0,B): text negatives for image b (weights_i2t);  blocks [B,2B): image negatives for text b (weights_t2i).
constexpr int HN_THREADS = 128;
__global__ void __launch_bounds__(HN_THREADS)
hard_negative_kernel(const float* __restrict__ image_feat, const float* __restrict__ text_feat, int B, int E,
                     const float* __restrict__ temp, const int64_t* __restrict__ idx, uint64_t seed, const uint64_t* __restrict__ salt,
                     float* __restrict__ w_i2t, float* __restrict__ w_t2i, int64_t* __restrict__ text_neg,
                     int64_t* __restrict__ image_neg) {
  extern __shared__ float sw[];  // [B] weights, then [E] the anchor row
  __shared__ float sh[32];
  float* anchor = sw + B;
  const bool t2i = blockIdx.x >= B;
  const int r = t2i ? blockIdx.x - B : blockIdx.x;
  const float* A = t2i ? text_feat : image_feat;
  const float* Bm = t2i ? image_feat : text_feat;
  for (int k = threadIdx.x; k < E; k += HN_THREADS) anchor[k] = A[(size_t)r * E + k];
  __syncthreads();
  const float inv_t = 1.f / *temp;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < B; j += HN_THREADS) {
    float d = 0.f;
    for (int k = 0; k < E; ++k) d = fmaf(anchor[k], Bm[(size_t)j * E + k], d);
    d *= inv_t;
    sw[j] = d;
    mx = fmaxf(mx, d);
  }
  mx = block_max<HN_THREADS>(mx, sh);
  float sum = 0.f;
  for (int j = threadIdx.x; j < B; j += HN_THREADS) {
    const float e = __expf(sw[j] - mx);
    sw[j] = e;
    sum += e;
  }
  sum = block_sum<HN_THREADS>(sum, sh);
  float* wout = t2i ? w_t2i : w_i2t;
  for (int j = threadIdx.x; j < B; j += HN_THREADS) {
    float w = sw[j] / sum + 1e-5f;
    const bool same = idx ? (idx[j] == idx[r]) : (j == r);
    if (same) w = 0.f;
    sw[j] = w;
    if (wout) wout[(size_t)r * B + j] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int j = 0; j < B; ++j) total += sw[j];
    const float u = hash_uniform(seed + *salt, (uint64_t)blockIdx.x) * total;
    float c = 0.f;
    int pick = -1, last = 0;
    for (int j = 0; j < B; ++j) {
      if (sw[j] > 0.f) {
        last = j;
        c += sw[j];
        if (pick < 0 && u < c) pick = j;
      }
    }
    if (pick < 0) pick = last;
    (t2i ? image_neg : text_neg)[r] = pick;
  }
}

int hard_negatives(const float* image_feat, const float* text_feat, int B, int E, const float* temp, const int64_t* idx,
                   uint64_t seed, float* w_i2t, float* w_t2i, int64_t* text_neg, int64_t* image_neg, cudaStream_t s) {
  hard_negative_kernel<<<2 * B, HN_THREADS, (size_t)(B + E) * sizeof(float), s>>>(image_feat, text_feat, B, E, temp, idx, seed, seed_salt_ptr(),
                                                                                w_i2t, w_t2i, text_neg, image_neg);
  count_launch();
  return (int)cudaGetLastError();
}

// ======================================================================================== VQ-KD codebook argmin
// ids[r] = argmin_n ( ||z_r||^2 + ||e_n||^2 - 2 z_r . e_n ), z l2-normalised over the 32 channels, first index on
// ties (norm_ema_quantizer.py:152-162).  Exact fp32 FMAs in the reference's operation order (sum, sum, -2*dot);
// neither `d` nor the one-hot matrix is materialised.  A CTA owns 128 z rows; the codebook streams through shared memory
// in tiles of 1024 codes.  A thread owns TWO rows and one of eight code slices of every tile (every lane of a warp reads the
// same code: broadcast loads, each feeding two dot products) with FOUR codes in flight — each dot product is still the
// sequential 32-step FMA chain of the oracle, but eight independent chains per thread and 16 warps per SM hide the FMA
// latency that bound the one-thread-per-row, one-code-at-a-time version (0.89 ms: one warp per scheduler waiting on a
// 32-deep dependent chain).  Candidates of the slices are merged by (distance, index), which is the first-index rule.
constexpr int VQ_DIM = 32;
constexpr int VQ_ROWS = 128;          // z rows per CTA
constexpr int VQ_RPT = 2;             // rows per thread: every code read from shared memory feeds two dot products
constexpr int VQ_SLICES = 8;          // code slices per row pair
constexpr int VQ_THREADS = VQ_ROWS / VQ_RPT * VQ_SLICES;   // 512
constexpr int VQ_TILE = 1024;  // codes per shared-memory tile (128 KB + 4 KB norms)
constexpr int VQ_ILP = 4;

__global__ void __launch_bounds__(VQ_THREADS)
vq_argmin_kernel(const float* __restrict__ z, const float* __restrict__ codebook, int64_t* __restrict__ ids, int R, int K) {
  extern __shared__ __align__(16) float vq_smem[];
  float* se = vq_smem;                     // [VQ_TILE][32]
  float* see = vq_smem + VQ_TILE * VQ_DIM; // [VQ_TILE]
  constexpr int PAIRS = VQ_ROWS / VQ_RPT;
  const int slice = threadIdx.x / PAIRS, lp = threadIdx.x % PAIRS;
  const int row0 = blockIdx.x * VQ_ROWS + lp * VQ_RPT;
  float zn[VQ_RPT][VQ_DIM];
  float zz[VQ_RPT];
#pragma unroll
  for (int t = 0; t < VQ_RPT; ++t) {
    const float* zr = z + (size_t)min(row0 + t, R - 1) * VQ_DIM;
    float nrm = 0.f;
#pragma unroll
    for (int k = 0; k < VQ_DIM; k += 4) {
      const float4 v = *(const float4*)(zr + k);
      zn[t][k] = v.x; zn[t][k + 1] = v.y; zn[t][k + 2] = v.z; zn[t][k + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < VQ_DIM; ++k) nrm = fmaf(zn[t][k], zn[t][k], nrm);
    const float denom = fmaxf(sqrtf(nrm), 1e-12f);  // F.normalize eps
#pragma unroll
    for (int k = 0; k < VQ_DIM; ++k) zn[t][k] = zn[t][k] / denom;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < VQ_DIM; ++k) acc = fmaf(zn[t][k], zn[t][k], acc);
    zz[t] = acc;
  }
  float best[VQ_RPT];
  int best_i[VQ_RPT];
#pragma unroll
  for (int t = 0; t < VQ_RPT; ++t) { best[t] = INFINITY; best_i[t] = 0; }
  for (int c0 = 0; c0 < K; c0 += VQ_TILE) {
    const int nc = min(VQ_TILE, K - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nc * (VQ_DIM / 4); i += VQ_THREADS)
      ((float4*)se)[i] = ((const float4*)(codebook + (size_t)c0 * VQ_DIM))[i];
    __syncthreads();
    for (int c = threadIdx.x; c < nc; c += VQ_THREADS) {
      float e2 = 0.f;
#pragma unroll
      for (int k = 0; k < VQ_DIM; ++k) e2 = fmaf(se[c * VQ_DIM + k], se[c * VQ_DIM + k], e2);
      see[c] = e2;
    }
    __syncthreads();
    // this slice's share of the tile, in ascending code order
    const int per = (nc + VQ_SLICES - 1) / VQ_SLICES;
    const int cb = slice * per, ce = min(nc, cb + per);
    int c = cb;
    for (; c + VQ_ILP <= ce; c += VQ_ILP) {
      float dot[VQ_RPT][VQ_ILP];
#pragma unroll
      for (int t = 0; t < VQ_RPT; ++t)
#pragma unroll
        for (int u = 0; u < VQ_ILP; ++u) dot[t][u] = 0.f;
#pragma unroll
      for (int k = 0; k < VQ_DIM / 4; ++k) {
#pragma unroll
        for (int u = 0; u < VQ_ILP; ++u) {
          const float4 e = ((const float4*)(se + (c + u) * VQ_DIM))[k];
#pragma unroll
          for (int t = 0; t < VQ_RPT; ++t) {
            dot[t][u] = fmaf(zn[t][4 * k], e.x, dot[t][u]);
            dot[t][u] = fmaf(zn[t][4 * k + 1], e.y, dot[t][u]);
            dot[t][u] = fmaf(zn[t][4 * k + 2], e.z, dot[t][u]);
            dot[t][u] = fmaf(zn[t][4 * k + 3], e.w, dot[t][u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < VQ_ILP; ++u) {
        const float e2 = see[c + u];
#pragma unroll
        for (int t = 0; t < VQ_RPT; ++t) {
          const float d = (zz[t] + e2) - 2.f * dot[t][u];
          if (d < best[t]) { best[t] = d; best_i[t] = c0 + c + u; }
        }
      }
    }
    for (; c < ce; ++c) {
      const float4* e4 = (const float4*)(se + c * VQ_DIM);
      float dot[VQ_RPT];
#pragma unroll
      for (int t = 0; t < VQ_RPT; ++t) dot[t] = 0.f;
#pragma unroll
      for (int k = 0; k < VQ_DIM / 4; ++k) {
        const float4 e = e4[k];
#pragma unroll
        for (int t = 0; t < VQ_RPT; ++t) {
          dot[t] = fmaf(zn[t][4 * k], e.x, dot[t]);
          dot[t] = fmaf(zn[t][4 * k + 1], e.y, dot[t]);
          dot[t] = fmaf(zn[t][4 * k + 2], e.z, dot[t]);
          dot[t] = fmaf(zn[t][4 * k + 3], e.w, dot[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < VQ_RPT; ++t) {
        const float d = (zz[t] + see[c]) - 2.f * dot[t];
        if (d < best[t]) { best[t] = d; best_i[t] = c0 + c; }
      }
    }
  }
  // merge the slices' candidates: smallest distance, then smallest index (= the first index among equal distances)
  __syncthreads();
  float* cd = vq_smem;                              // [VQ_SLICES][VQ_ROWS]
  int* ci = (int*)(vq_smem + VQ_SLICES * VQ_ROWS);  // [VQ_SLICES][VQ_ROWS]
#pragma unroll
  for (int t = 0; t < VQ_RPT; ++t) {
    cd[slice * VQ_ROWS + lp * VQ_RPT + t] = best[t];
    ci[slice * VQ_ROWS + lp * VQ_RPT + t] = best_i[t];
  }
  __syncthreads();
  if (threadIdx.x < VQ_ROWS) {
    const int lr = threadIdx.x, row = blockIdx.x * VQ_ROWS + lr;
    float b = cd[lr];
    int bi = ci[lr];
#pragma unroll
    for (int s2 = 1; s2 < VQ_SLICES; ++s2) {
      const float d = cd[s2 * VQ_ROWS + lr];
      const int i = ci[s2 * VQ_ROWS + lr];
      if (d < b || (d == b && i < bi)) { b = d; bi = i; }
    }
    if (row < R) ids[row] = bi;
  }
}

int vq_argmin(const float* z, const float* codebook, int64_t* ids, int R, int K, int C, cudaStream_t s) {
  if (C != VQ_DIM) { set_error("vq_argmin: codebook_dim must be 32 (got %d)", C); return XFM_ERR_BAD_ARG; }
  if (R <= 0) return 0;
  const size_t smem = (size_t)VQ_TILE * (VQ_DIM + 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(vq_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  vq_argmin_kernel<<<(R + VQ_ROWS - 1) / VQ_ROWS, VQ_THREADS, smem, s>>>(z, codebook, ids, R, K);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ CE host
int ce_fwd(const float* logits, int64_t ld, const int64_t* labels, int R, int V, float* row_loss, float* lse, float* loss,
           float* count, cudaStream_t s) {
  if (R <= 0) return 0;
  ce_fwd_kernel<<<R, CE_THREADS, 0, s>>>(logits, ld, labels, V, row_loss, lse);
  count_launch();
  ce_reduce_kernel<<<1, CE_THREADS, 0, s>>>(row_loss, labels, R, loss, count);
  count_launch();
  return (int)cudaGetLastError();
}
int ce_bwd(const float* logits, int64_t ld, const int64_t* labels, const float* lse, const float* count, const float* upstream,
           bf16* dlogits, int64_t ldd, int R, int V, cudaStream_t s) {
  if (R <= 0) return 0;
  ce_bwd_kernel<<<R, CE_THREADS, 0, s>>>(logits, ld, labels, lse, count, nullptr, upstream, dlogits, ldd, V);
  count_launch();
  return (int)cudaGetLastError();
}
int ce_bwd_rows(const float* logits, int64_t ld, const int64_t* labels, const float* lse, const float* row_scale,
                const float* upstream, bf16* dlogits, int64_t ldd, int R, int V, cudaStream_t s) {
  if (R <= 0) return 0;
  if (!row_scale) { set_error("ce_bwd_rows: row_scale is required"); return XFM_ERR_BAD_ARG; }
  ce_bwd_kernel<<<R, CE_THREADS, 0, s>>>(logits, ld, labels, lse, nullptr, row_scale, upstream, dlogits, ldd, V);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace xfm

// ======================================================================================== feature normalisation
// F.normalize(x, dim=-1) of the ITC projections (xfm.py:614-621): y = x / max(||x||, 1e-12); one warp per row.
namespace xfm {

__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm, int R, int E) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + (size_t)row * E;
  float s = 0.f;
  for (int k = lane; k < E; k += 32) s = fmaf(xr[k], xr[k], s);
  s = warp_sum(s);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int k = lane; k < E; k += 32) y[(size_t)row * E + k] = xr[k] * inv;
  if (lane == 0) inv_norm[row] = inv;
}
// dx = (dy - y * <y, dy>) * inv_norm
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ inv_norm,
                  float* __restrict__ dx, int R, int E) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* dr = dy + (size_t)row * E;
  const float* yr = y + (size_t)row * E;
  float s = 0.f;
  for (int k = lane; k < E; k += 32) s = fmaf(yr[k], dr[k], s);
  s = warp_sum(s);
  const float inv = inv_norm[row];
  for (int k = lane; k < E; k += 32) dx[(size_t)row * E + k] = (dr[k] - yr[k] * s) * inv;
}

int l2norm_fwd(const float* x, float* y, float* inv_norm, int R, int E, cudaStream_t s) {
  if (R <= 0) return 0;
  l2norm_fwd_kernel<<<(R + 7) / 8, 256, 0, s>>>(x, y, inv_norm, R, E);
  count_launch();
  return (int)cudaGetLastError();
}
int l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int R, int E, cudaStream_t s) {
  if (R <= 0) return 0;
  l2norm_bwd_kernel<<<(R + 7) / 8, 256, 0, s>>>(dy, y, inv_norm, dx, R, E);
  count_launch();
  return (int)cudaGetLastError();
}

// ======================================================================================== MIM (MSE variant)
// xfm.py:631-635: loss = mse(x[:,1:][mask], t[:,1:][mask]) (+ mse(x[:,0], t[:,0]) unless mim_cls_only).
// Forward and the gradient for an upstream of 1 in one pass over x / t (f32 [B, np+1, D]); rows outside the
// selection get a zero gradient.  *count = number of masked patches (device-side, no host sync).
__global__ void __launch_bounds__(256)
mask_count_kernel(const uint8_t* __restrict__ mask, size_t n, float* __restrict__ count) {
  __shared__ float sh[32];
  float c = 0.f;
  for (size_t i = threadIdx.x; i < n; i += 256) c += mask[i] ? 1.f : 0.f;
  c = block_sum<256>(c, sh);
  if (threadIdx.x == 0) *count = c;
}
__global__ void __launch_bounds__(256)
mim_mse_kernel(const float* __restrict__ x, const float* __restrict__ t, const uint8_t* __restrict__ mask,
               const float* __restrict__ count, int B, int np, int D, int with_cls, float* __restrict__ loss,
               float* __restrict__ dx) {
  __shared__ float sh[32];
  const int row = blockIdx.x;
  const int b = row / (np + 1), tok = row % (np + 1);
  float w = 0.f;
  if (tok == 0) w = with_cls ? 1.0f / ((float)B * (float)D) : 0.f;
  else if (mask[(size_t)b * np + tok - 1]) w = 1.0f / (*count * (float)D);
  float acc = 0.f;
  for (int c = threadIdx.x * 4; c < D; c += 256 * 4) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w != 0.f) {
      const float4 a = *(const float4*)(x + (size_t)row * D + c);
      const float4 r = *(const float4*)(t + (size_t)row * D + c);
      const float d0 = a.x - r.x, d1 = a.y - r.y, d2 = a.z - r.z, d3 = a.w - r.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      g = make_float4(2.f * w * d0, 2.f * w * d1, 2.f * w * d2, 2.f * w * d3);
    }
    *(float4*)(dx + (size_t)row * D + c) = g;
  }
  if (w != 0.f) {  // block-uniform
    acc = block_sum<256>(acc, sh);
    if (threadIdx.x == 0) atomicAdd(loss, acc * w);
  }
}
int mim_mse(const float* x, const float* t, const uint8_t* mask, int B, int np, int D, int with_cls, float* count,
            float* loss, float* dx, cudaStream_t s) {
  if (D & 3) { set_error("mim_mse: D must be a multiple of 4"); return XFM_ERR_BAD_ARG; }
  cudaMemsetAsync(loss, 0, sizeof(float), s);
  mask_count_kernel<<<1, 256, 0, s>>>(mask, (size_t)B * np, count);
  count_launch();
  mim_mse_kernel<<<B * (np + 1), 256, 0, s>>>(x, t, mask, count, B, np, D, with_cls, loss, dx);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace xfm
