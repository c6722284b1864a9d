// Region / bounding-box branch of the pre-training path: the per-sample gather + attention-weighted pooling
// of image tokens (models/beit2.py:468-475, models/xfm.py:574-597) and the box losses (models/xfm.py:815-840,
// models/box_ops.py).  All HBM-bound / tiny: coalesced float4 rows, one block per sample row.
#include "common.cuh"
#include "internal.h"

namespace xfm {

XFM_DEVINL void store_bf16x4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *(uint32_t*)&a;
  u.y = *(uint32_t*)&b;
  *(uint2*)p = u;
}

// out[b, 0, :]   = sum_j w[b, j] * y[idx[b], 1 + j, :] / sum_j w[b, j]      (w = image_atts[b, 1:], beit2.py:470-472)
// out[b, 1+j, :] = y[idx[b], 1 + j, :]                                       (x_bs, beit2.py:469)
// y: f32 [n_img, N, D] (token 0 = the full-image mean, unused here); idx int64 [bsz]; atts int64 [bsz, N].
__global__ void __launch_bounds__(256)
region_pool_fwd_kernel(const float* __restrict__ y, const int64_t* __restrict__ idx, const int64_t* __restrict__ atts,
                       float* __restrict__ out, bf16* __restrict__ out16, int N, int D) {
  const int b = blockIdx.x;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const float* src = y + (size_t)idx[b] * N * D;
  float* dst = out + (size_t)b * N * D;
  bf16* dst16 = out16 ? out16 + (size_t)b * N * D : nullptr;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float wsum = 0.f;
  for (int j = 1; j < N; ++j) {
    const float w = (float)atts[(size_t)b * N + j];
    const float4 v = *(const float4*)(src + (size_t)j * D + c);
    *(float4*)(dst + (size_t)j * D + c) = v;
    if (dst16) store_bf16x4(dst16 + (size_t)j * D + c, v);
    acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    wsum += w;
  }
  const float inv = 1.0f / wsum;
  acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  *(float4*)(dst + c) = acc;
  if (dst16) store_bf16x4(dst16 + c, acc);
}

// dy[idx[b], 1+j, :] += dout[b, 1+j, :] + w[b, j] / sum_j w[b, j] * dout[b, 0, :]   (several samples share an image: atomics)
__global__ void __launch_bounds__(256)
region_pool_bwd_kernel(const float* __restrict__ dout, const int64_t* __restrict__ idx, const int64_t* __restrict__ atts,
                       float* __restrict__ dy, int N, int D) {
  const int b = blockIdx.x;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  float wsum = 0.f;
  for (int j = 1; j < N; ++j) wsum += (float)atts[(size_t)b * N + j];
  const float inv = 1.0f / wsum;
  const float* g = dout + (size_t)b * N * D;
  float* dst = dy + (size_t)idx[b] * N * D;
  const float4 g0 = *(const float4*)(g + c);
  for (int j = 1; j < N; ++j) {
    const float w = (float)atts[(size_t)b * N + j] * inv;
    const float4 v = *(const float4*)(g + (size_t)j * D + c);
    float* p = dst + (size_t)j * D + c;
    atomicAdd(p + 0, v.x + w * g0.x);
    atomicAdd(p + 1, v.y + w * g0.y);
    atomicAdd(p + 2, v.z + w * g0.z);
    atomicAdd(p + 3, v.w + w * g0.w);
  }
}

int region_pool_fwd(const float* y, const int64_t* idx, const int64_t* atts, float* out, bf16_t* out16, int bsz, int N, int D,
                    cudaStream_t s) {
  if (D & 3 || bsz <= 0) { set_error("region_pool: D must be a multiple of 4, bsz > 0"); return XFM_ERR_BAD_ARG; }
  const int threads = D / 4 < 256 ? ((D / 4 + 31) / 32 * 32) : 256;
  dim3 grid(bsz, (D / 4 + threads - 1) / threads);
  region_pool_fwd_kernel<<<grid, threads, 0, s>>>(y, idx, atts, out, out16, N, D);
  count_launch();
  return (int)cudaGetLastError();
}
int region_pool_bwd(const float* dout, const int64_t* idx, const int64_t* atts, float* dy, int bsz, int N, int D,
                    cudaStream_t s) {
  if (D & 3 || bsz <= 0) { set_error("region_pool: D must be a multiple of 4, bsz > 0"); return XFM_ERR_BAD_ARG; }
  const int threads = D / 4 < 256 ? ((D / 4 + 31) / 32 * 32) : 256;
  dim3 grid(bsz, (D / 4 + threads - 1) / threads);
  region_pool_bwd_kernel<<<grid, threads, 0, s>>>(dout, idx, atts, dy, N, D);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- box head: sigmoid
__global__ void sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = 1.0f / (1.0f + __expf(-x[i]));
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dy[i] * y[i] * (1.0f - y[i]);
}
int sigmoid_fwd(const float* x, float* y, int n, cudaStream_t s) {
  if (n <= 0) return 0;
  sigmoid_fwd_kernel<<<(n + 255) / 256, 256, 0, s>>>(x, y, n);
  count_launch();
  return (int)cudaGetLastError();
}
int sigmoid_bwd(const float* dy, const float* y, float* dx, int n, cudaStream_t s) {
  if (n <= 0) return 0;
  sigmoid_bwd_kernel<<<(n + 255) / 256, 256, 0, s>>>(dy, y, dx, n);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- L1 + GIoU box loss
// coord, target: f32 [n, 4] (cx, cy, w, h); is_image: f32 [n] or null (1 = whole-image sample, excluded; xfm.py:833-838).
//   loss_bbox = sum_i keep_i * |coord_i - target_i|_1 / num_boxes,  loss_giou = sum_i keep_i * (1 - GIoU_ii) / num_boxes,
//   num_boxes = n or sum_i (1 - is_image_i).  If ANY box of either set is degenerate (x2 < x1 or y2 < y1) the GIoU loss is
//   zero for the whole batch (xfm.py:825-828) — decided on the device, no host sync.
// d_bbox / d_giou: gradients of the two losses wrt coord (for upstream gradients of 1).  One block.
constexpr int BOX_THREADS = 256;
__global__ void __launch_bounds__(BOX_THREADS)
bbox_loss_kernel(const float* __restrict__ coord, const float* __restrict__ target, const float* __restrict__ is_image, int n,
                 float* __restrict__ loss_bbox, float* __restrict__ loss_giou, float* __restrict__ d_bbox,
                 float* __restrict__ d_giou) {
  __shared__ float sh[3][BOX_THREADS / 32];
  __shared__ float tot[3];
  // pass 1: number of boxes, degenerate flag
  float cnt = 0.f;
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += BOX_THREADS) {
    cnt += is_image ? 1.0f - is_image[i] : 1.0f;
    const float4 c = ((const float4*)coord)[i], t = ((const float4*)target)[i];
    bad |= (c.z < 0.f) | (c.w < 0.f) | (t.z < 0.f) | (t.w < 0.f);   // x2 < x1  <=>  w < 0
  }
  const int any_bad = __syncthreads_or(bad);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int k = 0; k < BOX_THREADS / 32; ++k) r += sh[0][k];
    tot[0] = r;
  }
  __syncthreads();
  const float inv_boxes = 1.0f / tot[0];
  float l1 = 0.f, lg = 0.f;
  for (int i = threadIdx.x; i < n; i += BOX_THREADS) {
    const float keep = is_image ? 1.0f - is_image[i] : 1.0f;
    const float4 c = ((const float4*)coord)[i], t = ((const float4*)target)[i];
    const float cc[4] = {c.x, c.y, c.z, c.w}, tt[4] = {t.x, t.y, t.z, t.w};
    float g1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = cc[k] - tt[k];
      l1 += keep * fabsf(d);
      g1[k] = keep * inv_boxes * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    ((float4*)d_bbox)[i] = make_float4(g1[0], g1[1], g1[2], g1[3]);
    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!any_bad) {
      const float x1 = c.x - 0.5f * c.z, y1 = c.y - 0.5f * c.w, x2 = c.x + 0.5f * c.z, y2 = c.y + 0.5f * c.w;
      const float u1 = t.x - 0.5f * t.z, v1 = t.y - 0.5f * t.w, u2 = t.x + 0.5f * t.z, v2 = t.y + 0.5f * t.w;
      const float a1 = (x2 - x1) * (y2 - y1), a2 = (u2 - u1) * (v2 - v1);
      const float iw = fmaxf(fminf(x2, u2) - fmaxf(x1, u1), 0.f), ih = fmaxf(fminf(y2, v2) - fmaxf(y1, v1), 0.f);
      const float inter = iw * ih, uni = a1 + a2 - inter;
      const float cw = fmaxf(fmaxf(x2, u2) - fminf(x1, u1), 0.f), ch = fmaxf(fmaxf(y2, v2) - fminf(y1, v1), 0.f);
      const float ac = cw * ch;
      const float giou = inter / uni - (ac - uni) / ac;
      lg += keep * (1.0f - giou);
      // d/d(x1, y1, x2, y2) of inter, union and the enclosing area
      float di[4] = {0.f, 0.f, 0.f, 0.f}, da[4] = {0.f, 0.f, 0.f, 0.f};
      if (iw > 0.f && ih > 0.f) {
        if (x1 > u1) di[0] = -ih;
        if (y1 > v1) di[1] = -iw;
        if (x2 < u2) di[2] = ih;
        if (y2 < v2) di[3] = iw;
      }
      if (cw > 0.f && ch > 0.f) {
        if (x1 < u1) da[0] = -ch;
        if (y1 < v1) da[1] = -cw;
        if (x2 > u2) da[2] = ch;
        if (y2 > v2) da[3] = cw;
      }
      const float d1[4] = {-(y2 - y1), -(x2 - x1), (y2 - y1), (x2 - x1)};
      float dl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float du = d1[k] - di[k];
        const float dg = (di[k] * uni - inter * du) / (uni * uni) + (du * ac - uni * da[k]) / (ac * ac);
        dl[k] = -dg * keep * inv_boxes;
      }
      gg = make_float4(dl[0] + dl[2], dl[1] + dl[3], 0.5f * (dl[2] - dl[0]), 0.5f * (dl[3] - dl[1]));
    }
    ((float4*)d_giou)[i] = gg;
  }
  l1 = warp_sum(l1);
  lg = warp_sum(lg);
  if ((threadIdx.x & 31) == 0) { sh[1][threadIdx.x >> 5] = l1; sh[2][threadIdx.x >> 5] = lg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float r1 = 0.f, r2 = 0.f;
    for (int k = 0; k < BOX_THREADS / 32; ++k) { r1 += sh[1][k]; r2 += sh[2][k]; }
    *loss_bbox = r1 * inv_boxes;
    *loss_giou = r2 * inv_boxes;
  }
}

int bbox_loss(const float* coord, const float* target, const float* is_image, int n, float* loss_bbox, float* loss_giou,
              float* d_bbox, float* d_giou, cudaStream_t s) {
  if (n <= 0) { set_error("bbox_loss: n must be > 0"); return XFM_ERR_BAD_ARG; }
  bbox_loss_kernel<<<1, BOX_THREADS, 0, s>>>(coord, target, is_image, n, loss_bbox, loss_giou, d_bbox, d_giou);
  count_launch();
  return (int)cudaGetLastError();
}

// out = a * sa[0] + b * sb[0]   (two upstream loss gradients applied to two prepared gradients)
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ sa, const float* __restrict__ b,
                             const float* __restrict__ sb, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * (sa ? *sa : 0.f) + b[i] * (sb ? *sb : 0.f);
}
int axpby_scalars(const float* a, const float* sa, const float* b, const float* sb, float* out, int n, cudaStream_t s) {
  if (n <= 0) return 0;
  axpby_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, sa, b, sb, out, n);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace xfm
