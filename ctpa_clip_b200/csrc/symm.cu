// Latent exchange over NVLink peer memory, fused with the global-batch logits (SURVEY §8(e), K14 + C1).
//
// The reference trains with local-batch InfoNCE under DDP (CTCLIPTrainer.py:213-217, ct_clip.py:845-878); the
// north star asks for the global batch: every rank needs all ranks' l2-normalised latents [T_hat; I_hat]. Instead of
// an NCCL all-gather followed by the logits kernel, ONE kernel does both:
//   * CTA (p, 0) PUSHES this rank's 2*b*d floats into rank p's symmetric buffer with 16-byte peer stores and then
//     releases flag[rank] on p (st.release.sys);
//   * CTA (p, q) waits (ld.acquire.sys) only for the two flags it depends on — the text rows of rank p and the image
//     rows of rank q in ITS OWN buffer — and computes that b x b block of L = exp(tau) T I^T, so blocks of ranks that
//     arrived early are computed while later ranks' latents are still in flight.
// The row/column log-sum-exp and the gradient kernels of loss.cu follow on the same stream and read the gathered
// latents in place. Flags carry the caller's step counter (no reset pass); two buffer parities make the push of step
// s+1 safe while a slow peer still reads step s (a push of step s+2 needs that peer's flag of step s+1, which it
// only releases after its step-s kernels, in stream order).
//
// Symmetric buffers are the one place this library owns device memory (ctclip_b200.h "Ownership"): they must come
// from plain cudaMalloc to be exportable through CUDA IPC, so ctclip_symm_alloc / _free wrap it explicitly.
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {

constexpr int kMaxWorld = 32;
constexpr int kFlagStride = 8;  // one flag per 32 bytes
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000 * 1000 * 1000;

struct PeerTable {
  float* buf[kMaxWorld];
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {  // L2 (coherence point for peer stores), never L1
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// buffer layout (floats): [0, 2*kMaxWorld*kFlagStride) flags as u32: flag[parity][rank] at (parity*kMaxWorld+rank)*8;
// then per parity: T [world*b][d], I [world*b][d]
__host__ __device__ inline size_t header_floats() { return (size_t)2 * kMaxWorld * kFlagStride; }
__host__ __device__ inline size_t parity_floats(int world, int b, int d) { return (size_t)2 * world * b * d; }

// grid (world, world), 256 threads.
__global__ void __launch_bounds__(256)
latent_exchange_logits_kernel(PeerTable peers, const float* __restrict__ t_hat, const float* __restrict__ i_hat,
                              const float* __restrict__ tau, int b, int d, int rank, int world, unsigned step,
                              float* __restrict__ L) {
  const int p = blockIdx.x, q = blockIdx.y;
  const int par = step & 1u;
  const int B = world * b;
  const size_t par_off = header_floats() + (size_t)par * parity_floats(world, b, d);
  if (q == 0) {
    // ---- push this rank's rows into rank p's buffer (p == rank: plain local stores), then publish
    float* dstT = peers.buf[p] + par_off + (size_t)rank * b * d;
    float* dstI = dstT + (size_t)B * d;
    const int n4 = b * d / 4;
    for (int k = threadIdx.x; k < n4; k += blockDim.x) {
      reinterpret_cast<float4*>(dstT)[k] = reinterpret_cast<const float4*>(t_hat)[k];
      reinterpret_cast<float4*>(dstI)[k] = reinterpret_cast<const float4*>(i_hat)[k];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      st_release_sys(reinterpret_cast<unsigned*>(peers.buf[p]) + (par * kMaxWorld + rank) * kFlagStride, step);
  }
  // ---- wait for the text rows of rank p and the image rows of rank q in OUR buffer
  __shared__ int timed_out;
  if (threadIdx.x == 0) {
    const unsigned* flags = reinterpret_cast<const unsigned*>(peers.buf[rank]) + par * kMaxWorld * kFlagStride;
    const unsigned long long t0 = globaltimer_ns();
    int bad = 0;
    while (ld_acquire_sys(flags + p * kFlagStride) != step || ld_acquire_sys(flags + q * kFlagStride) != step) {
      if (globaltimer_ns() - t0 > kSpinTimeoutNs) { bad = 1; break; }
      __nanosleep(64);
    }
    timed_out = bad;
  }
  __syncthreads();
  const float* T = peers.buf[rank] + par_off;
  const float* I = T + (size_t)B * d;
  const float et = __expf(*tau);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int pr = warp; pr < b * b; pr += nwarps) {
    const int i = p * b + pr / b, j = q * b + pr % b;
    float s = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      const float4 a = ld_cg_f4(T + (size_t)i * d + k);
      const float4 c = ld_cg_f4(I + (size_t)j * d + k);
      s += (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w);
    }
    s = warp_sum(s);
    // a peer that never arrived poisons the loss instead of hanging the box
    if (lane == 0) L[(size_t)i * B + j] = timed_out ? __int_as_float(0x7fc00000) : et * s;
  }
}

}  // namespace

namespace ctclip {
// loss.cu: row/column log-sum-exp + gradient of the local rows on already formed logits
int clip_lse_grad_launch(const float* T, const float* I, const float* tau, int B, int d, int row0, int rows_local,
                         float* work, float* loss, float* dT, float* dI, float* dtau, cudaStream_t s);
}  // namespace ctclip

extern "C" size_t ctclip_symm_latent_bytes(int b_local, int d, int world) {
  if (b_local <= 0 || d <= 0 || world <= 0 || world > kMaxWorld) return 0;
  return (header_floats() + 2 * parity_floats(world, b_local, d)) * sizeof(float);
}

extern "C" int ctclip_symm_alloc(size_t bytes, void** ptr) {
  if (ptr == nullptr || bytes == 0) return ctclip::fail(CTCLIP_E_SHAPE, "symm_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "symm_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(*ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "symm_alloc: cudaMemset: %s", cudaGetErrorString(e));
  return CTCLIP_OK;
}

extern "C" int ctclip_symm_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "symm_free: %s", cudaGetErrorString(e));
  return CTCLIP_OK;
}

extern "C" int ctclip_symm_export(void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "symm_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(handle64, &h, 64);
  return CTCLIP_OK;
}

extern "C" int ctclip_symm_import(const unsigned char* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return ctclip::fail(CTCLIP_E_CUDA, "symm_import: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  }
  return CTCLIP_OK;
}

extern "C" int ctclip_symm_unimport(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "symm_unimport: %s", cudaGetErrorString(e));
  return CTCLIP_OK;
}

extern "C" int ctclip_clip_loss_allgather(const float* t_hat, const float* i_hat, const float* tau, int b_local, int d,
                                          int rank, int world, void* const* host_peer_bufs, unsigned step, float* work,
                                          float* loss, float* dT, float* dI, float* dtau, void* stream) {
  if (b_local <= 0 || d <= 0 || d % 4) return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss_allgather: bad shape b=%d d=%d", b_local, d);
  if (world <= 0 || world > kMaxWorld || rank < 0 || rank >= world)
    return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss_allgather: bad rank %d / world %d (max %d)", rank, world, kMaxWorld);
  if (step == 0) return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss_allgather: step counter starts at 1 (flags are zero-initialised)");
  if ((long long)world * b_local > 8192) return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss_allgather: global batch too large");
  if (host_peer_bufs == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss_allgather: no peer table");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  PeerTable peers;
  for (int r = 0; r < kMaxWorld; ++r) peers.buf[r] = r < world ? static_cast<float*>(host_peer_bufs[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    if (peers.buf[r] == nullptr || (reinterpret_cast<uintptr_t>(peers.buf[r]) & 15))
      return ctclip::fail(CTCLIP_E_ALIGN, "clip_loss_allgather: peer buffer %d missing or not 16-byte aligned", r);
  if ((reinterpret_cast<uintptr_t>(t_hat) | reinterpret_cast<uintptr_t>(i_hat)) & 15)
    return ctclip::fail(CTCLIP_E_ALIGN, "clip_loss_allgather: latents must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int B = world * b_local;
  float* L = work;
  latent_exchange_logits_kernel<<<dim3(world, world), 256, 0, s>>>(peers, t_hat, i_hat, tau, b_local, d, rank, world, step, L);
  rc = ctclip::check_launch("latent_exchange_logits");
  if (rc) return rc;
  const size_t par_off = header_floats() + (size_t)(step & 1u) * parity_floats(world, b_local, d);
  const float* T = peers.buf[rank] + par_off;
  const float* I = T + (size_t)B * d;
  return ctclip::clip_lse_grad_launch(T, I, tau, B, d, rank * b_local, b_local, work, loss, dT, dI, dtau, s);
}
