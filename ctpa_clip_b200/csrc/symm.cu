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
#include <cstdlib>

namespace {

constexpr int kMaxWorld = 32;
constexpr int kFlagStride = 8;  // one flag per 32 bytes
// How long a CTA waits for a peer's flag before it gives up: default 600 s (torch.distributed's NCCL default; the
// reference's accelerate process group uses 36000 s, CTCLIPTrainer.py:212), CTCLIP_PEER_TIMEOUT_S or
// ctclip_symm_set_timeout_ms override it. Giving up poisons the logits with NaN AND raises the caller's status word, so
// the host can tell a dead peer from a numerical problem; the optimiser kernel skips any update whose gradient norm is
// not finite (optim.cu), so a timeout never becomes a training update.
unsigned long long g_timeout_ns = 0;   // 0 = not initialised
unsigned long long spin_timeout_ns() {
  if (g_timeout_ns == 0) {
    const char* e = getenv("CTCLIP_PEER_TIMEOUT_S");
    double sec = e != nullptr ? atof(e) : 600.0;
    if (!(sec > 0.0)) sec = 600.0;
    g_timeout_ns = (unsigned long long)(sec * 1e9);
  }
  return g_timeout_ns;
}

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

// grid (world, world), 256 threads. Emulation (rank < 0, grid (world, world, world), cooperative launch so that all CTAs
// are co-resident): blockIdx.z plays the rank — all `world` symmetric buffers live on ONE device, t_hat / i_hat / L hold
// every rank's slice back to back — which exercises the push / flag / wait protocol on a single GPU in a single launch.
__global__ void __launch_bounds__(256)
latent_exchange_logits_kernel(PeerTable peers, const float* __restrict__ t_hat, const float* __restrict__ i_hat,
                              const float* __restrict__ tau, int b, int d, int rank, int world, unsigned step,
                              float* __restrict__ L, unsigned long long timeout_ns, int* __restrict__ status) {
  const int p = blockIdx.x, q = blockIdx.y;
  const int par = step & 1u;
  const int B = world * b;
  if (rank < 0) {
    rank = blockIdx.z;
    t_hat += (size_t)rank * b * d;
    i_hat += (size_t)rank * b * d;
    L += (size_t)rank * B * B;
  }
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
      if (globaltimer_ns() - t0 > timeout_ns) { bad = 1; break; }
      __nanosleep(64);
    }
    timed_out = bad;
    if (bad && status != nullptr) atomicOr(status, 1);
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
    // a peer that never arrived poisons the loss (and raised *status above) instead of hanging the box
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

extern "C" int ctclip_symm_set_timeout_ms(unsigned long long ms) {
  if (ms == 0) return ctclip::fail(CTCLIP_E_SHAPE, "symm_set_timeout_ms: the timeout must be positive");
  g_timeout_ns = ms * 1000000ull;
  return CTCLIP_OK;
}

namespace {
int check_exchange_args(const char* who, const float* t_hat, const float* i_hat, int b_local, int d, int world,
                        void* const* host_peer_bufs, unsigned step, PeerTable* peers) {
  if (b_local <= 0 || d <= 0 || d % 4) return ctclip::fail(CTCLIP_E_SHAPE, "%s: bad shape b=%d d=%d", who, b_local, d);
  if (world <= 0 || world > kMaxWorld) return ctclip::fail(CTCLIP_E_SHAPE, "%s: bad world %d (max %d)", who, world, kMaxWorld);
  if (step == 0) return ctclip::fail(CTCLIP_E_SHAPE, "%s: step counter starts at 1 (flags are zero-initialised)", who);
  if ((long long)world * b_local > 8192) return ctclip::fail(CTCLIP_E_SHAPE, "%s: global batch too large", who);
  if (host_peer_bufs == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "%s: no peer table", who);
  for (int r = 0; r < kMaxWorld; ++r) peers->buf[r] = r < world ? static_cast<float*>(host_peer_bufs[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    if (peers->buf[r] == nullptr || (reinterpret_cast<uintptr_t>(peers->buf[r]) & 15))
      return ctclip::fail(CTCLIP_E_ALIGN, "%s: peer buffer %d missing or not 16-byte aligned", who, r);
  if ((reinterpret_cast<uintptr_t>(t_hat) | reinterpret_cast<uintptr_t>(i_hat)) & 15)
    return ctclip::fail(CTCLIP_E_ALIGN, "%s: latents must be 16-byte aligned", who);
  return ctclip::require_sm100();
}
}  // namespace

extern "C" int ctclip_clip_loss_allgather(const float* t_hat, const float* i_hat, const float* tau, int b_local, int d,
                                          int rank, int world, void* const* host_peer_bufs, unsigned step, float* work,
                                          float* loss, float* dT, float* dI, float* dtau, int* status, void* stream) {
  if (rank < 0 || rank >= world) return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss_allgather: bad rank %d / world %d", rank, world);
  PeerTable peers;
  int rc = check_exchange_args("clip_loss_allgather", t_hat, i_hat, b_local, d, world, host_peer_bufs, step, &peers);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int B = world * b_local;
  float* L = work;
  latent_exchange_logits_kernel<<<dim3(world, world), 256, 0, s>>>(peers, t_hat, i_hat, tau, b_local, d, rank, world, step, L,
                                                                  spin_timeout_ns(), status);
  rc = ctclip::check_launch("latent_exchange_logits");
  if (rc) return rc;
  const size_t par_off = header_floats() + (size_t)(step & 1u) * parity_floats(world, b_local, d);
  const float* T = peers.buf[rank] + par_off;
  const float* I = T + (size_t)B * d;
  return ctclip::clip_lse_grad_launch(T, I, tau, B, d, rank * b_local, b_local, work, loss, dT, dI, dtau, s);
}

// All `world` ranks of the exchange on ONE device in ONE cooperative launch (bring-up / single-GPU test of the push / flag /
// wait protocol; B200_PROFILING.md: mutually waiting kernels must never be separate launches on one GPU). bufs[r] are
// `world` symmetric buffers of this device; t_hat_all / i_hat_all fp32 [world*b_local][d]; per-rank outputs back to back:
// work_all [world][B*B + 2*B + d], loss_all [world], dT_all / dI_all [world*b_local][d], dtau_all [world].
extern "C" int ctclip_clip_loss_allgather_emulated(const float* t_hat_all, const float* i_hat_all, const float* tau,
                                                   int b_local, int d, int world, void* const* host_bufs, unsigned step,
                                                   float* work_all, float* loss_all, float* dT_all, float* dI_all,
                                                   float* dtau_all, int* status, void* stream) {
  PeerTable peers;
  int rc = check_exchange_args("clip_loss_allgather_emulated", t_hat_all, i_hat_all, b_local, d, world, host_bufs, step, &peers);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int B = world * b_local;
  const size_t work_stride = (size_t)B * B + 2 * (size_t)B + d;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, latent_exchange_logits_kernel, 256, 0);
  if ((long long)world * world * world > (long long)per_sm * ctclip::sm_count())
    return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss_allgather_emulated: %d^3 CTAs cannot be co-resident", world);
  // the per-rank logits land at work_all + rank * B * B inside the kernel; the per-rank work areas are work_stride apart,
  // so the logits are computed into a packed scratch at the head of work_all[0..] and copied out below
  int rank = -1;
  unsigned long long timeout = spin_timeout_ns();
  float* Lall = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&Lall), (size_t)world * B * B * sizeof(float), s);
  if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "clip_loss_allgather_emulated: scratch: %s", cudaGetErrorString(e));
  void* args[] = {&peers, &t_hat_all, &i_hat_all, &tau, &b_local, &d, &rank, &world, &step, &Lall, &timeout, &status};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(latent_exchange_logits_kernel), dim3(world, world, world), dim3(256),
                                  args, 0, s);
  if (e != cudaSuccess) {
    cudaFreeAsync(Lall, s);
    return ctclip::fail(CTCLIP_E_CUDA, "clip_loss_allgather_emulated: cooperative launch: %s", cudaGetErrorString(e));
  }
  ctclip::count_launch();
  const size_t par_off = header_floats() + (size_t)(step & 1u) * parity_floats(world, b_local, d);
  for (int r = 0; r < world && rc == 0; ++r) {
    float* work = work_all + r * work_stride;
    cudaMemcpyAsync(work, Lall + (size_t)r * B * B, (size_t)B * B * sizeof(float), cudaMemcpyDeviceToDevice, s);
    const float* T = peers.buf[r] + par_off;
    const float* I = T + (size_t)B * d;
    rc = ctclip::clip_lse_grad_launch(T, I, tau, B, d, r * b_local, b_local, work, loss_all + r,
                                      dT_all + (size_t)r * b_local * d, dI_all + (size_t)r * b_local * d, dtau_all + r, s);
  }
  cudaFreeAsync(Lall, s);
  return rc;
}
