// Flat-buffer optimizer step for the data-parallel path (replaces accelerators/ddp_accelerator.py:89-98 +
// optim.py:4-50): global grad-norm, clip, HF-style AdamW (transformers.optimization.AdamW: bias-corrected step
// size, eps added to sqrt(v), decoupled weight decay applied after the Adam update) and the refresh of the bf16
// shadow the tcgen05 GEMMs read — one HBM pass over (P, G, m, v) instead of ~750 per-tensor launches.
//
// Parameters live in one fp32 buffer split in 64-element chunks; chunk_group[c] selects the hyper-parameter group
// (0..3 = {decay, no-decay} x {lr, lr*mult}; 255 = frozen / no gradient this step -> untouched, like a None grad).
#include "common.cuh"
#include "internal.h"

namespace xfm {

constexpr int OPT_THREADS = 256;

// out[0] = sum g^2 over chunks whose group != 255.  DETERMINISTIC: every block writes its partial sum, the block that
// arrives last adds the partials in index order.  (An atomicAdd per block made the clip factor differ in the last bits
// between data-parallel ranks that hold identical gradients, and the replicas drifted apart by ulps per step.)
constexpr int SUMSQ_MAX_BLOCKS = 4096;
__device__ float g_sumsq_partials[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_arrived = 0;

__global__ void __launch_bounds__(OPT_THREADS)
sumsq_kernel(const float* __restrict__ g, const uint8_t* __restrict__ chunk_group, size_t nchunks, float* __restrict__ out) {
  __shared__ float sh[OPT_THREADS / 32];
  float acc = 0.f;
  const size_t nvec = nchunks * 16;  // float4 per chunk = 16
  for (size_t i = (size_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < nvec; i += (size_t)gridDim.x * OPT_THREADS) {
    if (chunk_group && chunk_group[i >> 4] == 255) continue;
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
      *out = r;
      g_sumsq_arrived = 0;
    }
  }
}

struct AdamArgs {
  float lr[4], wd[4];
  float beta1, beta2, eps, bias_c1, bias_c2;  // bias_c1 = 1 - beta1^t, bias_c2 = 1 - beta2^t
  float max_norm, grad_mul;                   // max_norm <= 0: no clipping; grad_mul: 1/world for DDP averaging
};

__global__ void __launch_bounds__(OPT_THREADS)
adamw_flat_kernel(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ M, float* __restrict__ V,
                  bf16* __restrict__ S, const uint8_t* __restrict__ chunk_group, size_t nchunks,
                  const float* __restrict__ sumsq, float* __restrict__ norm_out, const AdamArgs a) {
  float clip = a.grad_mul;
  if (sumsq) {
    const float total = sqrtf(*sumsq) * a.grad_mul;
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
    if (a.max_norm > 0.f) clip *= fminf(1.0f, a.max_norm / (total + 1e-6f));  // torch.nn.utils.clip_grad_norm_
  }
  const size_t nvec = nchunks * 16;
  for (size_t i = (size_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < nvec; i += (size_t)gridDim.x * OPT_THREADS) {
    const int grp = chunk_group[i >> 4];
    if (grp == 255) continue;
    const float lr = a.lr[grp], wd = a.wd[grp];
    const float step = lr * sqrtf(a.bias_c2) / a.bias_c1;
    float4 p = ((float4*)P)[i];
    const float4 g4 = ((const float4*)G)[i];
    float4 m = ((float4*)M)[i], v = ((float4*)V)[i];
    float* pp = (float*)&p; const float* gg = (const float*)&g4; float* mm = (float*)&m; float* vv = (float*)&v;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = gg[k] * clip;
      mm[k] = mm[k] * a.beta1 + g * (1.f - a.beta1);
      vv[k] = vv[k] * a.beta2 + g * g * (1.f - a.beta2);
      float x = pp[k] - step * mm[k] / (sqrtf(vv[k]) + a.eps);
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

int grad_sumsq(const float* g, const uint8_t* chunk_group, size_t nchunks, float* out, cudaStream_t s) {
  int grid = num_sms() * 8;
  if (grid > SUMSQ_MAX_BLOCKS) grid = SUMSQ_MAX_BLOCKS;
  sumsq_kernel<<<grid, OPT_THREADS, 0, s>>>(g, chunk_group, nchunks, out);
  count_launch();
  return (int)cudaGetLastError();
}

int adamw_flat(float* P, const float* G, float* M, float* V, bf16_t* S, const uint8_t* chunk_group, size_t nchunks,
               const float* sumsq, float* norm_out, const xfm_adamw_params* hp, cudaStream_t s) {
  if (!hp || hp->step < 1) { set_error("adamw: step must be >= 1"); return XFM_ERR_BAD_ARG; }
  AdamArgs a;
  for (int i = 0; i < 4; ++i) { a.lr[i] = hp->lr[i]; a.wd[i] = hp->weight_decay[i]; }
  a.beta1 = hp->beta1; a.beta2 = hp->beta2; a.eps = hp->eps;
  a.bias_c1 = hp->correct_bias ? 1.0f - powf(hp->beta1, (float)hp->step) : 1.0f;
  a.bias_c2 = hp->correct_bias ? 1.0f - powf(hp->beta2, (float)hp->step) : 1.0f;
  a.max_norm = hp->max_grad_norm; a.grad_mul = hp->grad_mul;
  const int grid = num_sms() * 8;
  adamw_flat_kernel<<<grid, OPT_THREADS, 0, s>>>(P, G, M, V, S, chunk_group, nchunks, sumsq, norm_out, a);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace xfm
