// Host-side runtime of libctclip_sm100.so: error reporting, device checks, TMA descriptor encoding.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "ctclip_internal.h"

namespace {
thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
}  // namespace

namespace ctclip {

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int require_sm100() {
  static int cached[64];
  static bool init = false;
  if (!init) {
    memset(cached, 0, sizeof(cached));
    init = true;
  }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(CTCLIP_E_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  if (dev >= 0 && dev < 64 && cached[dev] == 1) return CTCLIP_OK;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return fail(CTCLIP_E_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (major != 10) return fail(CTCLIP_E_ARCH, "device %d is sm_%d0, this library only runs on sm_100a (B200)", dev, major);
  if (dev >= 0 && dev < 64) cached[dev] = 1;
  return CTCLIP_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, int rank, void* base, const cuuint64_t* dims,
                const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* elem_strides,
                CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(CTCLIP_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUresult r = fn(map, dtype, static_cast<cuuint32_t>(rank), base, dims, strides_bytes, box, elem_strides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CTCLIP_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CTCLIP_OK;
}

}  // namespace ctclip

extern "C" int ctclip_version(void) { return 100; }

extern "C" int ctclip_last_error(char* buf, size_t n) {
  if (buf == nullptr || n == 0) return (int)strlen(g_err);
  strncpy(buf, g_err, n - 1);
  buf[n - 1] = 0;
  return (int)strlen(buf);
}

extern "C" long long ctclip_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
