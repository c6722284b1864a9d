// xfm_b200 — host-side internals shared by the .cu translation units (not part of the C-ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/xfm_b200.h"

namespace xfm {

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn get_tensor_map_encoder();
void set_error(const char* fmt, ...);
int num_sms();
void count_launch(int n = 1);

int gemm_bf16(const xfm_gemm_params* p, cudaStream_t stream);

}  // namespace xfm
