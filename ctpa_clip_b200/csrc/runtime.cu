// Host-side runtime of libctclip_sm100.so: error reporting, device checks, TMA descriptor encoding.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

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

// SMs the persistent kernels size their grids for: all of them, or the budget set by ctclip_set_sm_budget / CTCLIP_SM_BUDGET
// (data-parallel training: a static persistent schedule over 148 CTAs runs a second wave for every SM a concurrent NCCL
// kernel holds; sizing the grids a few SMs short removes that tail).
static std::atomic<int> g_sm_budget{-1};
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int n = (dev >= 0 && dev < 64) ? cached[dev] : 0;
  if (n <= 0) {
    n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < 64) cached[dev] = n;
  }
  int b = g_sm_budget.load(std::memory_order_relaxed);
  if (b < 0) {
    const char* e = getenv("CTCLIP_SM_BUDGET");
    b = e != nullptr ? atoi(e) : 0;
    g_sm_budget.store(b, std::memory_order_relaxed);
  }
  return (b > 0 && b < n) ? b : n;
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

extern "C" int ctclip_set_sm_budget(int sms) {
  if (sms < 0) return ctclip::fail(CTCLIP_E_SHAPE, "set_sm_budget: negative budget");
  ctclip::g_sm_budget.store(sms, std::memory_order_relaxed);   // 0 = all SMs
  return CTCLIP_OK;
}

extern "C" int ctclip_last_error(char* buf, size_t n) {
  if (buf == nullptr || n == 0) return (int)strlen(g_err);
  strncpy(buf, g_err, n - 1);
  buf[n - 1] = 0;
  return (int)strlen(buf);
}

extern "C" long long ctclip_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Caller-owned workspace sizes (bytes) of the entry points that take one; dims as documented per op in ctclip_b200.h.
extern "C" long long ctclip_workspace_bytes(const char* op, const long long* dims, int ndims) {
  if (op == nullptr || (ndims > 0 && dims == nullptr)) return -1;
  auto need = [&](int n) { return ndims == n; };
  const long long f = (long long)sizeof(float);
  if (!strcmp(op, "clip_loss") && need(2)) return (dims[0] * dims[0] + 2 * dims[0] + dims[1]) * f;            // (B, d)
  if (!strcmp(op, "clip_loss_allgather") && need(3)) {                                                        // (b_local, d, world)
    const long long B = dims[0] * dims[2];
    return (B * B + 2 * B + dims[1]) * f;
  }
  if ((!strcmp(op, "cpb_table_fwd") || !strcmp(op, "cpb_table_bwd")) && need(3)) {                            // (h, w, dim)
    const long long R = (2 * dims[0] - 1) * (2 * dims[1] - 1);
    return (!strcmp(op, "cpb_table_fwd") ? 2 * R + 2 * R * dims[2] : 2 * R * dims[2]) * f;
  }
  if (!strcmp(op, "bert_attn_bwd") && need(3)) return dims[0] * dims[1] * dims[2] * f;                        // (batch, heads, seq_len)
  if (!strcmp(op, "prep_resample") && need(0)) return 8192 * f;                                               // exact HU table
  if (!strcmp(op, "sumsq") && need(0)) return 1024 * f;                                                       // per-block partials
  ctclip::fail(CTCLIP_E_SHAPE, "workspace_bytes: unknown op '%s' or wrong number of dims (%d)", op, ndims);
  return -1;
}
