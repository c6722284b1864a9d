// Flat-buffer optimizer step for the data-parallel path (replaces accelerators/ddp_accelerator.py:89-98 +
// optim.py:4-50): global grad-norm, clip, HF-style AdamW (transformers.optimization.AdamW: bias-corrected step
// size, eps added to sqrt(v), decoupled weight decay applied after the Adam update) and the refresh of the bf16
// shadow the tcgen05 GEMMs read — one HBM pass over (P, G, m, v) instead of ~750 per-tensor launches.
//
// Parameters live in one fp32 buffer split in 64-element chunks.  chunk_seg[c] (static) is the parameter segment a
// chunk belongs to (-1: frozen / padding, never touched); seg_group[s] (per step) is that parameter's hyper-parameter group
// (0..3 = {decay, no-decay} x {lr, lr*mult}) or 255 when it received no gradient since the last zero_grad (the
// reference's optimizer skips grad-None parameters, and its per-parameter `step` only advances when it is updated:
// seg_step[s] is that counter, kept on the device so that nothing in the step depends on a host-side count).
// Hyper-parameters are read from a 16-float DEVICE block (xfm_adamw_hparams) — an LR scheduler only rewrites that block.
#include "common.cuh"
#include "internal.h"

namespace xfm {

constexpr int OPT_THREADS = 256;

// out[0] (+)= sum g^2 over chunks of live segments.  DETERMINISTIC: every block writes its partial sum, the block that
// arrives last adds the partials in index order.  (An atomicAdd per block made the clip factor differ in the last bits
// between data-parallel ranks that hold identical gradients, and the replicas drifted apart by ulps per step.)
// Block 0 also advances the per-segment step counters and derives the bias corrections the update kernel reads.
constexpr int SUMSQ_MAX_BLOCKS = 4096;
__device__ float g_sumsq_partials[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_arrived = 0;

__global__ void __launch_bounds__(OPT_THREADS)
sumsq_kernel(const float* __restrict__ g, const int32_t* __restrict__ chunk_seg, const uint8_t* __restrict__ seg_group,
             size_t nchunks, int32_t* __restrict__ seg_step, float2* __restrict__ seg_bc, int nseg,
             const float* __restrict__ hp, float* __restrict__ out, int accumulate) {
  __shared__ float sh[OPT_THREADS / 32];
  if (blockIdx.x == 0 && seg_step) {
    const float b1 = hp[8], b2 = hp[9];
    const bool correct = hp[13] != 0.f;
    for (int s = threadIdx.x; s < nseg; s += OPT_THREADS) {
      if (seg_group[s] == 255) continue;
      const int t = ++seg_step[s];
      seg_bc[s] = correct ? make_float2(1.0f - powf(b1, (float)t), 1.0f - powf(b2, (float)t)) : make_float2(1.f, 1.f);
    }
  }
  float acc = 0.f;
  const size_t nvec = nchunks * 16;  // float4 per chunk = 16
  const size_t stride = (size_t)gridDim.x * OPT_THREADS;
  size_t i = (size_t)blockIdx.x * OPT_THREADS + threadIdx.x;
  // four independent (segment lookup -> live test -> 16-byte load) chains in flight per thread: one chain per iteration
  // left the kernel at 4.0 TB/s (0.62 of the copy rate)
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    int seg[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) seg[k] = chunk_seg[(i + k * stride) >> 4];
    bool live[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) live[k] = seg[k] >= 0 && seg_group[seg[k] < 0 ? 0 : seg[k]] != 255;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = live[k] ? __ldg((const float4*)g + i + k * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  }
  for (; i < nvec; i += stride) {
    const int seg = chunk_seg[i >> 4];
    if (seg < 0 || seg_group[seg] == 255) continue;
    const float4 v = ((const float4*)g)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < OPT_THREADS / 32 ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) g_sumsq_partials[blockIdx.x] = r;
  }
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&g_sumsq_arrived, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += OPT_THREADS) t += ((volatile float*)g_sumsq_partials)[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < OPT_THREADS / 32 ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) {
      *out = accumulate ? *out + r : r;
      g_sumsq_arrived = 0;
    }
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_flat_kernel(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ M, float* __restrict__ V,
                  bf16* __restrict__ S, const int32_t* __restrict__ chunk_seg, const uint8_t* __restrict__ seg_group,
                  const float2* __restrict__ seg_bc, size_t nchunks, const float* __restrict__ sumsq,
                  float* __restrict__ norm_out, const float* __restrict__ hp) {
  // hp: lr[4] | wd[4] | beta1 beta2 eps max_norm | grad_mul correct_bias - -
  const float beta1 = hp[8], beta2 = hp[9], eps = hp[10], max_norm = hp[11], grad_mul = hp[12];
  float clip = grad_mul;
  if (sumsq) {
    const float total = sqrtf(*sumsq) * grad_mul;
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
    if (max_norm > 0.f) clip *= fminf(1.0f, max_norm / (total + 1e-6f));  // torch.nn.utils.clip_grad_norm_
  }
  const size_t nvec = nchunks * 16;
  for (size_t i = (size_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < nvec; i += (size_t)gridDim.x * OPT_THREADS) {
    const int seg = chunk_seg[i >> 4];
    if (seg < 0) continue;
    const int grp = seg_group[seg];
    if (grp == 255) continue;
    const float lr = hp[grp], wd = hp[4 + grp];
    const float2 bc = seg_bc[seg];
    const float step = lr * sqrtf(bc.y) / bc.x;
    float4 p = ((float4*)P)[i];
    const float4 g4 = ((const float4*)G)[i];
    float4 m = ((float4*)M)[i], v = ((float4*)V)[i];
    float* pp = (float*)&p; const float* gg = (const float*)&g4; float* mm = (float*)&m; float* vv = (float*)&v;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = gg[k] * clip;
      mm[k] = mm[k] * beta1 + g * (1.f - beta1);
      vv[k] = vv[k] * beta2 + g * g * (1.f - beta2);
      float x = pp[k] - step * mm[k] / (sqrtf(vv[k]) + eps);
      if (wd > 0.f) x -= lr * wd * x;
      pp[k] = x;
    }
    ((float4*)P)[i] = p;
    ((float4*)M)[i] = m;
    ((float4*)V)[i] = v;
    if (S) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
      uint2 u;
      u.x = *(uint32_t*)&lo; u.y = *(uint32_t*)&hi;
      ((uint2*)S)[i] = u;
    }
  }
}

int grad_sumsq(const float* g, const int32_t* chunk_seg, const uint8_t* seg_group, size_t nchunks, int32_t* seg_step,
               float* seg_bc, int nseg, const float* hp, float* out, int accumulate, cudaStream_t s) {
  if (!g || !chunk_seg || !seg_group || !out || (seg_step && (!seg_bc || !hp))) {
    set_error("grad_sumsq: null argument");
    return XFM_ERR_BAD_ARG;
  }
  int grid = num_sms() * 8;
  if (grid > SUMSQ_MAX_BLOCKS) grid = SUMSQ_MAX_BLOCKS;
  sumsq_kernel<<<grid, OPT_THREADS, 0, s>>>(g, chunk_seg, seg_group, nchunks, seg_step, (float2*)seg_bc, nseg, hp, out,
                                            accumulate);
  count_launch();
  return (int)cudaGetLastError();
}

int adamw_flat(float* P, const float* G, float* M, float* V, bf16_t* S, const int32_t* chunk_seg, const uint8_t* seg_group,
               const float* seg_bc, size_t nchunks, const float* sumsq, float* norm_out, const float* hp, cudaStream_t s) {
  if (!P || !G || !M || !V || !chunk_seg || !seg_group || !seg_bc || !hp) {
    set_error("adamw: null argument");
    return XFM_ERR_BAD_ARG;
  }
  const int grid = num_sms() * 8;
  adamw_flat_kernel<<<grid, OPT_THREADS, 0, s>>>(P, G, M, V, S, chunk_seg, seg_group, (const float2*)seg_bc, nchunks, sumsq,
                                                 norm_out, hp);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace xfm
