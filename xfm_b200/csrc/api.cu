// xfm_b200 — extern "C" entry points (the C-ABI declared in include/xfm_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "internal.h"

namespace xfm {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static TensorMapEncodeFn g_encode = nullptr;
static int g_num_sms = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
TensorMapEncodeFn get_tensor_map_encoder() { return g_encode; }
int num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }

}  // namespace xfm

using namespace xfm;

extern "C" {

int xfm_version(void) { return 1; }

int xfm_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return XFM_ERR_NO_DRIVER;
    }
    g_encode = (TensorMapEncodeFn)fn;
  }
  return 0;
}

int64_t xfm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* xfm_last_error(void) { return g_err; }

int xfm_gemm_bf16(const xfm_gemm_params* p, void* stream) { return gemm_bf16(p, (cudaStream_t)stream); }

}  // extern "C"
