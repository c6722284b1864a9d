// Device half of the batch feeder (loader side): the tail of the reference's image transform —
// transforms.ToTensor() + transforms.Normalize(mean, std) (dataset/__init__.py:26-35) and, optionally, the
// RandomHorizontalFlip of the same Compose — applied on the GPU to the uint8 HWC crops a decode worker produces, so
// that a step's images cross PCIe as 1 byte per channel value instead of 4.  HBM-bound: 3 B read + 12 B written per pixel.
//
// Arithmetic is the reference's, operation by operation, with IEEE round-to-nearest divisions (no reciprocal
// multiplication): ToTensor = float(u8) / 255 ; Normalize = (x - mean) / std.  The result is bit-identical to the
// CPU transform on the same crop (tests/test_feed_gpu.py).
#include "common.cuh"
#include "internal.h"

namespace xfm {

struct NormConst {
  float mean[3];
  float stdv[3];
};

XFM_DEVINL float to_norm(uint32_t u, float mean, float stdv) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), mean), stdv);
}

// in  u8  [B, H, W, 3]; out f32 [B, 3, H, W]; flip u8 [B] or null (1 = mirror the row).  A byte has 256 values: every CTA first
// tabulates the transform per channel (3 x 256 floats in shared memory, two IEEE divisions per entry — 6 per thread instead of
// 24 per quad, which had made the first version issue-bound at 0.45 of the HBM rate), then walks quads of 4 consecutive output
// pixels of one row in a grid-stride loop: 12 contiguous input bytes (three 32-bit loads; a warp reads 384 contiguous bytes),
// 12 table look-ups, one float4 per colour plane (a warp writes 3 x 512 contiguous bytes).  W % 4 == 0.
__global__ void __launch_bounds__(256)
image_u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, const uint8_t* __restrict__ flip,
                       NormConst nc, int H, int W, size_t total_quads) {
  __shared__ float lut[3][256];
#pragma unroll
  for (int c = 0; c < 3; ++c) lut[c][threadIdx.x] = to_norm(threadIdx.x, nc.mean[c], nc.stdv[c]);
  __syncthreads();
  const int wq = W >> 2;
  const size_t plane = (size_t)H * W;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += (size_t)gridDim.x * blockDim.x) {
    const int xq = (int)(q % wq);
    const size_t row = q / wq;                 // b * H + y
    const size_t b = row / H;
    const int y = (int)(row - b * H);
    const bool mirror = flip != nullptr && flip[b] != 0;
    const int x_in = mirror ? W - 4 - 4 * xq : 4 * xq;
    const uint32_t* src = (const uint32_t*)(in + (row * W + x_in) * 3);   // 12-byte aligned quads: (row*W + x_in) % 4 == 0
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    // bytes: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
    float4 r = make_float4(lut[0][w0 & 255u], lut[0][w0 >> 24], lut[0][(w1 >> 16) & 255u], lut[0][(w2 >> 8) & 255u]);
    float4 g = make_float4(lut[1][(w0 >> 8) & 255u], lut[1][w1 & 255u], lut[1][w1 >> 24], lut[1][(w2 >> 16) & 255u]);
    float4 bl = make_float4(lut[2][(w0 >> 16) & 255u], lut[2][(w1 >> 8) & 255u], lut[2][w2 & 255u], lut[2][w2 >> 24]);
    if (mirror) {
      r = make_float4(r.w, r.z, r.y, r.x);
      g = make_float4(g.w, g.z, g.y, g.x);
      bl = make_float4(bl.w, bl.z, bl.y, bl.x);
    }
    float* dst = out + b * 3 * plane + (size_t)y * W + 4 * xq;
    *(float4*)dst = r;
    *(float4*)(dst + plane) = g;
    *(float4*)(dst + 2 * plane) = bl;
  }
}

int image_u8_to_f32(const uint8_t* in, float* out, const uint8_t* flip, int B, int H, int W, const float* mean, const float* stdv,
                    cudaStream_t s) {
  if (B < 0 || H <= 0 || W <= 0 || (W & 3)) { set_error("image_u8_to_f32: need B >= 0, H > 0, W > 0 and W %% 4 == 0"); return XFM_ERR_BAD_ARG; }
  if (B == 0) return 0;
  if (!in || !out || !mean || !stdv) { set_error("image_u8_to_f32: null pointer"); return XFM_ERR_BAD_ARG; }
  if (((uintptr_t)in & 3) || ((uintptr_t)out & 15)) { set_error("image_u8_to_f32: in must be 4-byte and out 16-byte aligned"); return XFM_ERR_BAD_ARG; }
  for (int c = 0; c < 3; ++c)
    if (!(stdv[c] != 0.0f)) { set_error("image_u8_to_f32: std evaluated to zero"); return XFM_ERR_BAD_ARG; }
  NormConst nc;
  for (int c = 0; c < 3; ++c) { nc.mean[c] = mean[c]; nc.stdv[c] = stdv[c]; }
  const size_t quads = (size_t)B * H * (W >> 2);
  const size_t want = (quads + 255) / 256, cap = (size_t)num_sms() * 8;   // 8 resident CTAs of 256 threads per SM
  const unsigned blocks = (unsigned)(want < cap ? want : cap);
  image_u8_to_f32_kernel<<<blocks, 256, 0, s>>>(in, out, flip, nc, H, W, quads);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Crop + bicubic resize of uint8 images: what `RandomResizedCrop(res, interpolation=BICUBIC)`, `Resize((res, res), BICUBIC)`
// (dataset/__init__.py:28-30,63-67) and `image.crop(...)` + `resize(image, [res, res], BICUBIC)` (dataset/pretrain_dataset.py:
// 470-483) do through torchvision -> Pillow (third party; Pillow's ImagingResample, src/libImaging/Resample.c): a separable
// convolution in 8.22 fixed point, horizontal pass first into a uint8 intermediate, then the vertical pass, each output
// = clip8((2^21 + sum_k coef[k] * pixel[k]) >> 22).  The coefficient tables come from the host (xfm_b200/feed.py restates
// Pillow's precompute_coeffs / normalize_coeffs_8bpc in float64); the kernels do the integer arithmetic, so the result is
// bit-identical to Pillow's (tests/test_feed_gpu.py compares with PIL itself).
//
// Images are ragged: `src` is one packed byte buffer; desc[b] = {byte offset of image b, its width in pixels, crop x0, crop y0,
// crop width, crop height, byte offset of its rows in `tmp`, 0}.
struct ResizeDesc {
  int64_t src_off, src_w, x0, y0, cw, ch, tmp_off, pad;
};

XFM_DEVINL uint8_t clip8_fixed(int v) {
  v >>= 22;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[b][row][x][c] = horizontal pass of crop row `row`; one thread per output pixel (3 channels), grid (x blocks, rows, B).
__global__ void __launch_bounds__(128)
resize_h_kernel(const uint8_t* __restrict__ src, const ResizeDesc* __restrict__ desc, const int32_t* __restrict__ hb,
                const int32_t* __restrict__ hk, int KH, uint8_t* __restrict__ tmp, int OW) {
  const int b = blockIdx.z, row = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
  const ResizeDesc d = desc[b];
  if (row >= d.ch || x >= OW) return;
  const int xmin = hb[((size_t)b * OW + x) * 2], n = hb[((size_t)b * OW + x) * 2 + 1];
  const int32_t* k = hk + ((size_t)b * OW + x) * KH;
  const uint8_t* p = src + d.src_off + ((d.y0 + row) * d.src_w + d.x0 + xmin) * 3;
  int r = 1 << 21, g = 1 << 21, bl = 1 << 21;
  for (int i = 0; i < n; ++i) {
    const int w = __ldg(k + i);
    r += w * p[3 * i]; g += w * p[3 * i + 1]; bl += w * p[3 * i + 2];
  }
  uint8_t* o = tmp + d.tmp_off + ((size_t)row * OW + x) * 3;
  o[0] = clip8_fixed(r); o[1] = clip8_fixed(g); o[2] = clip8_fixed(bl);
}

// out[b][y][x][c] = vertical pass over tmp; one thread per output BYTE (x * 3 + c): fully coalesced rows.
__global__ void __launch_bounds__(128)
resize_v_kernel(const uint8_t* __restrict__ tmp, const ResizeDesc* __restrict__ desc, const int32_t* __restrict__ vb,
                const int32_t* __restrict__ vk, int KV, uint8_t* __restrict__ out, int OH, int OW) {
  const int b = blockIdx.z, y = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= OW * 3) return;
  const ResizeDesc d = desc[b];
  const int ymin = vb[((size_t)b * OH + y) * 2], n = vb[((size_t)b * OH + y) * 2 + 1];
  const int32_t* k = vk + ((size_t)b * OH + y) * KV;
  const uint8_t* p = tmp + d.tmp_off + (size_t)ymin * OW * 3 + j;
  int acc = 1 << 21;
  for (int i = 0; i < n; ++i) acc += __ldg(k + i) * p[(size_t)i * OW * 3];
  out[((size_t)b * OH + y) * OW * 3 + j] = clip8_fixed(acc);
}

int resize_bicubic_u8(const uint8_t* src, const int64_t* desc, const int32_t* hb, const int32_t* hk, int KH, const int32_t* vb,
                      const int32_t* vk, int KV, uint8_t* tmp, uint8_t* out, int B, int max_rows, int OH, int OW, cudaStream_t s) {
  if (B < 0 || OH <= 0 || OW <= 0 || KH <= 0 || KV <= 0 || max_rows <= 0 || max_rows > 65535 || OH > 65535 || B > 65535) {
    set_error("resize_bicubic_u8: need 0 <= B <= 65535, 0 < OH, max_rows <= 65535, OW > 0, KH > 0, KV > 0");
    return XFM_ERR_BAD_ARG;
  }
  if (B == 0) return 0;
  if (!src || !desc || !hb || !hk || !vb || !vk || !tmp || !out) { set_error("resize_bicubic_u8: null pointer"); return XFM_ERR_BAD_ARG; }
  static_assert(sizeof(ResizeDesc) == 8 * sizeof(int64_t), "desc layout");
  resize_h_kernel<<<dim3((OW + 127) / 128, max_rows, B), 128, 0, s>>>(src, (const ResizeDesc*)desc, hb, hk, KH, tmp, OW);
  resize_v_kernel<<<dim3((OW * 3 + 127) / 128, OH, B), 128, 0, s>>>(tmp, (const ResizeDesc*)desc, vb, vk, KV, out, OH, OW);
  count_launch(2);
  return (int)cudaGetLastError();
}

// Tap tables on the device: Pillow's precompute_coeffs (box = the whole crop) + normalize_coeffs_8bpc for the bicubic filter
// (a = -0.5), one thread per output index and axis.  float64 with explicitly rounded operations (__dmul_rn / __dadd_rn /
// __ddiv_rn: no FMA contraction), in Pillow's operation order, so the integers are the ones Pillow computes on the host; the
// taps are evaluated twice (sum, then normalise) instead of being stored.  ~B * (OW + OH) threads of a few dozen FP64
// operations each: negligible even at Blackwell's FP64 rate, and it takes 22 ms of numpy per 96-image batch off the host.
XFM_DEVINL double bicubic_tap(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

__global__ void __launch_bounds__(128)
resize_taps_kernel(const ResizeDesc* __restrict__ desc, int32_t* __restrict__ hb, int32_t* __restrict__ hk, int KH,
                   int32_t* __restrict__ vb, int32_t* __restrict__ vk, int KV, int OH, int OW) {
  const int b = blockIdx.z, axis = blockIdx.y, xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int out_size = axis ? OH : OW, K = axis ? KV : KH;
  if (xx >= out_size) return;
  const int in_size = (int)(axis ? desc[b].ch : desc[b].cw);
  int32_t* bounds = (axis ? vb : hb) + ((size_t)b * out_size + xx) * 2;
  int32_t* taps = (axis ? vk : hk) + ((size_t)b * out_size + xx) * K;
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, fs);
  const double center = __dmul_rn((double)xx + 0.5, scale);
  const double ss = __ddiv_rn(1.0, fs);
  int first = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (first < 0) first = 0;
  int count = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (count > in_size) count = in_size;
  count -= first;
  if (count > K) count = K;          // cannot happen when the caller sized K = ceil(support) * 2 + 1
  double total = 0.0;
  for (int x = 0; x < count; ++x)
    total = __dadd_rn(total, bicubic_tap(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + first), center), 0.5), ss)));
  for (int x = 0; x < K; ++x) {
    int v = 0;
    if (x < count) {
      double w = bicubic_tap(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + first), center), 0.5), ss));
      if (total != 0.0) w = __ddiv_rn(w, total);
      const double q = __dmul_rn(w, 4194304.0);
      v = w < 0.0 ? (int)__dadd_rn(-0.5, q) : (int)__dadd_rn(0.5, q);
    }
    taps[x] = v;
  }
  bounds[0] = first;
  bounds[1] = count;
}

int resize_taps(const int64_t* desc, int32_t* hb, int32_t* hk, int KH, int32_t* vb, int32_t* vk, int KV, int B, int OH, int OW,
                cudaStream_t s) {
  if (B < 0 || B > 65535 || OH <= 0 || OW <= 0 || KH <= 0 || KV <= 0) { set_error("resize_taps: bad sizes"); return XFM_ERR_BAD_ARG; }
  if (B == 0) return 0;
  if (!desc || !hb || !hk || !vb || !vk) { set_error("resize_taps: null pointer"); return XFM_ERR_BAD_ARG; }
  const int n = OH > OW ? OH : OW;
  resize_taps_kernel<<<dim3((n + 127) / 128, 2, B), 128, 0, s>>>((const ResizeDesc*)desc, hb, hk, KH, vb, vk, KV, OH, OW);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace xfm
