// Memory-bound row-wise kernels of the CT-CLIP hot path: LayerNorm forward/backward, GEGLU forward/backward,
// fp32->bf16 casts. One warp per token row, 128-bit loads/stores, warp-shuffle reductions.
// Reference operators: attention.py:28-35 (gamma-only LayerNorm), :39-52 (nn.LayerNorm + GEGLU), ctvit.py:173.
#include "ptx.cuh"
#include "ctclip_internal.h"
#include <cstdlib>

namespace {

constexpr int kMaxVecAll = 8;  // float4 per lane -> dim <= 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------ LayerNorm fwd
// y = (x - mean) * rstd * gamma (+ beta). Optional outputs: bf16 normalised, bf16 raw copy of x, fp32 normalised.
template <int kMaxVec>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, long long rows, int dim, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y_bf16,
                     __nv_bfloat16* __restrict__ raw_bf16, float* __restrict__ y_f32) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = dim >> 2;
  for (long long row = warp_global; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
    float4 v[kMaxVec];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        v[j] = xr[i];
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    const float mean = warp_sum(s) / dim;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / dim + eps);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[i];
        float4 o;
        o.x = (v[j].x - mean) * rstd * g.x;
        o.y = (v[j].y - mean) * rstd * g.y;
        o.z = (v[j].z - mean) * rstd * g.z;
        o.w = (v[j].w - mean) * rstd * g.w;
        if (beta != nullptr) {
          const float4 b = reinterpret_cast<const float4*>(beta)[i];
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        if (y_f32 != nullptr) reinterpret_cast<float4*>(y_f32 + row * dim)[i] = o;
        if (y_bf16 != nullptr) {
          uint2 p;
          p.x = ptx::pack_bf16(o.x, o.y);
          p.y = ptx::pack_bf16(o.z, o.w);
          reinterpret_cast<uint2*>(y_bf16 + row * dim)[i] = p;
        }
        if (raw_bf16 != nullptr) {
          uint2 p;
          p.x = ptx::pack_bf16(v[j].x, v[j].y);
          p.y = ptx::pack_bf16(v[j].z, v[j].w);
          reinterpret_cast<uint2*>(raw_bf16 + row * dim)[i] = p;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm bwd
// dx_out = (add_in ? add_in : 0) + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy  (fp32 atomics, one per column per block)
template <int kMaxVec, bool DY_BF16, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
layernorm_bwd_kernel(const void* __restrict__ dy_v, const float* __restrict__ x, long long rows, int dim,
                     const float* __restrict__ gamma, float eps, const float* __restrict__ add_in,
                     float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int l2_ahead_on) {
  extern __shared__ float red[];  // [2][dim]
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long warp_global = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * warps_per_block;
  const int nvec = dim >> 2;
  for (int i = threadIdx.x; i < 2 * dim; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float4 acc_g[kMaxVec], acc_b[kMaxVec];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) acc_g[j] = acc_b[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  const bool l2_ahead = DY_BF16 && (dim == 512) && l2_ahead_on;   // measured: 181 -> 173 us with a bf16 dy, slower with an fp32 dy   // 2 KB fp32 rows = 16 lines: one line per lane of a half warp
  for (long long row = warp_global; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
    if (l2_ahead) {
      // No registers to spare for deeper software pipelining (3 CTAs / SM), so the lines this warp needs later are pulled
      // into L2 now: the residual-gradient row consumed after the reduction round, and the next row's x / dy / add_in.
      const long long nrow = row + nwarps;
      const int ln = lane & 15;
      if (add_in != nullptr) {
        const float* pa = (lane < 16) ? add_in + row * dim : (nrow < rows ? add_in + nrow * dim : nullptr);
        if (pa != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + ln * 32));
      }
      if (nrow < rows) {
        if (lane < 16) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(x + nrow * dim + ln * 32));
        } else if (DY_BF16) {
          if (ln < 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const __nv_bfloat16*>(dy_v) + nrow * dim + ln * 64));
        } else {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float*>(dy_v) + nrow * dim + ln * 32));
        }
      }
    }
    float4 v[kMaxVec], d[kMaxVec];
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        v[j] = xr[i];
        if (DY_BF16) {   // upstream gradient straight from a bf16 GEMM epilogue
          const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_v) + row * dim)[i];
          d[j] = make_float4(ptx::bf16_lo(u.x), ptx::bf16_hi(u.x), ptx::bf16_lo(u.y), ptx::bf16_hi(u.y));
        } else {
          d[j] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_v) + row * dim)[i];
        }
      }
    }
    // ONE reduction round for the four row sums. The statistics are taken about x0 = the row's first element (a sample of
    // the row, so |mean - x0| ~ std): sum(x - x0), sum((x - x0)^2) give mean and variance without cancellation, and
    // sum(g xhat) = rstd (sum(g (x - x0)) - (mean - x0) sum(g)),  g = dy * gamma.
    const float x0 = __shfl_sync(0xffffffffu, v[0].x, 0);
    float s1 = 0.f, s2 = 0.f, sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        v[j].x -= x0; v[j].y -= x0; v[j].z -= x0; v[j].w -= x0;
        s1 += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        s2 += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
        const float gx = d[j].x * gm.x, gy = d[j].y * gm.y, gz = d[j].z * gm.z, gw = d[j].w * gm.w;
        sg += (gx + gy) + (gz + gw);
        sgx += (gx * v[j].x + gy * v[j].y) + (gz * v[j].z + gw * v[j].w);
      }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, m);
      s2 += __shfl_xor_sync(0xffffffffu, s2, m);
      sg += __shfl_xor_sync(0xffffffffu, sg, m);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, m);
    }
    asm volatile("" ::: "memory");   // gamma is re-read (L1) below instead of living in 16 registers across the reduction
    const float inv_dim = 1.f / dim;
    const float m0 = s1 * inv_dim;                                     // mean - x0
    const float rstd = rsqrtf(fmaxf(s2 * inv_dim - m0 * m0, 0.f) + eps);
    const float mg = sg * inv_dim, mgx = rstd * (sgx - m0 * sg) * inv_dim;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        // v <- xhat ; accumulate parameter grads ; dx
        v[j].x = (v[j].x - m0) * rstd; v[j].y = (v[j].y - m0) * rstd;
        v[j].z = (v[j].z - m0) * rstd; v[j].w = (v[j].w - m0) * rstd;
        acc_g[j].x += d[j].x * v[j].x; acc_g[j].y += d[j].y * v[j].y;
        acc_g[j].z += d[j].z * v[j].z; acc_g[j].w += d[j].w * v[j].w;
        acc_b[j].x += d[j].x; acc_b[j].y += d[j].y; acc_b[j].z += d[j].z; acc_b[j].w += d[j].w;
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
        float4 o;
        o.x = rstd * (d[j].x * gm.x - mg - v[j].x * mgx);
        o.y = rstd * (d[j].y * gm.y - mg - v[j].y * mgx);
        o.z = rstd * (d[j].z * gm.z - mg - v[j].z * mgx);
        o.w = rstd * (d[j].w * gm.w - mg - v[j].w * mgx);
        if (add_in != nullptr) {
          const float4 a = reinterpret_cast<const float4*>(add_in + row * dim)[i];
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        if (dx_out != nullptr) reinterpret_cast<float4*>(dx_out + row * dim)[i] = o;
        if (dx_bf16 != nullptr) {
          uint2 p;
          p.x = ptx::pack_bf16(o.x, o.y);
          p.y = ptx::pack_bf16(o.z, o.w);
          reinterpret_cast<uint2*>(dx_bf16 + row * dim)[i] = p;
        }
      }
    }
  }
  // block reduction of the parameter gradients through shared memory, then one atomic per column
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      atomicAdd(&red[4 * i + 0], acc_g[j].x); atomicAdd(&red[4 * i + 1], acc_g[j].y);
      atomicAdd(&red[4 * i + 2], acc_g[j].z); atomicAdd(&red[4 * i + 3], acc_g[j].w);
      atomicAdd(&red[dim + 4 * i + 0], acc_b[j].x); atomicAdd(&red[dim + 4 * i + 1], acc_b[j].y);
      atomicAdd(&red[dim + 4 * i + 2], acc_b[j].z); atomicAdd(&red[dim + 4 * i + 3], acc_b[j].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    if (dgamma != nullptr) atomicAdd(dgamma + i, red[i]);
    if (dbeta != nullptr) atomicAdd(dbeta + i, red[dim + i]);
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm bwd, dim = 512
// The image tower's case ([110592, 512] rows, 44 % of the LayerNorm time of a step). Same arithmetic, in the same order, as
// layernorm_bwd_kernel<4, ...> (bit-identical dx); the difference is how the rows reach the lanes: every warp owns a two-stage
// shared-memory ring and requests the NEXT row's x, dy and add_in with cp.async (16-byte pieces, no registers held) while it
// works on the current one, so 16 warps / SM keep ~160 KB of reads in flight instead of what fits the register file.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool DY_BF16>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd512_kernel(const void* __restrict__ dy_v, const float* __restrict__ x, long long rows,
                        const float* __restrict__ gamma, float eps, const float* __restrict__ add_in,
                        float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                        float* __restrict__ dbeta) {
  constexpr int DIM = 512, NV = 4;
  constexpr int DYB = DY_BF16 ? 1024 : 2048;
  constexpr int STAGE = 4096 + DYB;                 // [x 2 KB | add_in 2 KB | dy]
  extern __shared__ __align__(16) uint8_t sm512[];
  float* red = reinterpret_cast<float*>(sm512);     // [2][512]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* ring = sm512 + 2 * DIM * 4 + warp * 2 * STAGE;
  const long long warp_global = (long long)blockIdx.x * 8 + warp;
  const long long nwarps = (long long)gridDim.x * 8;
  for (int i = threadIdx.x; i < 2 * DIM; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float4 acc_g[NV], acc_b[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc_g[j] = acc_b[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto issue = [&](long long row, int st) {
    uint8_t* b = ring + st * STAGE;
    const float* xr = x + row * DIM;
#pragma unroll
    for (int j = 0; j < NV; ++j) cp_async16(b + (lane + 32 * j) * 16, xr + (lane + 32 * j) * 4);
    if (add_in != nullptr) {
      const float* ar = add_in + row * DIM;
#pragma unroll
      for (int j = 0; j < NV; ++j) cp_async16(b + 2048 + (lane + 32 * j) * 16, ar + (lane + 32 * j) * 4);
    }
    if (DY_BF16) {
      const __nv_bfloat16* dr = reinterpret_cast<const __nv_bfloat16*>(dy_v) + row * DIM;
#pragma unroll
      for (int j = 0; j < 2; ++j) cp_async16(b + 4096 + (lane + 32 * j) * 16, dr + (lane + 32 * j) * 8);
    } else {
      const float* dr = reinterpret_cast<const float*>(dy_v) + row * DIM;
#pragma unroll
      for (int j = 0; j < NV; ++j) cp_async16(b + 4096 + (lane + 32 * j) * 16, dr + (lane + 32 * j) * 4);
    }
  };

  int st = 0;
  if (warp_global < rows) issue(warp_global, 0);
  cp_async_commit();
  for (long long row = warp_global; row < rows; row += nwarps, st ^= 1) {
    if (row + nwarps < rows) issue(row + nwarps, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();      // this row's group has landed (the newest one may still be in flight)
    __syncwarp();            // ... for every lane's pieces
    const uint8_t* b = ring + st * STAGE;
    float4 v[NV], d[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = lane + 32 * j;
      v[j] = *reinterpret_cast<const float4*>(b + i * 16);
      if (DY_BF16) {
        const uint2 u = *reinterpret_cast<const uint2*>(b + 4096 + i * 8);
        d[j] = make_float4(ptx::bf16_lo(u.x), ptx::bf16_hi(u.x), ptx::bf16_lo(u.y), ptx::bf16_hi(u.y));
      } else {
        d[j] = *reinterpret_cast<const float4*>(b + 4096 + i * 16);
      }
    }
    const float x0 = __shfl_sync(0xffffffffu, v[0].x, 0);
    float s1 = 0.f, s2 = 0.f, sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = lane + 32 * j;
      v[j].x -= x0; v[j].y -= x0; v[j].z -= x0; v[j].w -= x0;
      s1 += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      s2 += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      const float gx = d[j].x * gm.x, gy = d[j].y * gm.y, gz = d[j].z * gm.z, gw = d[j].w * gm.w;
      sg += (gx + gy) + (gz + gw);
      sgx += (gx * v[j].x + gy * v[j].y) + (gz * v[j].z + gw * v[j].w);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, m);
      s2 += __shfl_xor_sync(0xffffffffu, s2, m);
      sg += __shfl_xor_sync(0xffffffffu, sg, m);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, m);
    }
    asm volatile("" ::: "memory");
    const float inv_dim = 1.f / DIM;
    const float m0 = s1 * inv_dim;                                     // mean - x0
    const float rstd = rsqrtf(fmaxf(s2 * inv_dim - m0 * m0, 0.f) + eps);
    const float mg = sg * inv_dim, mgx = rstd * (sgx - m0 * sg) * inv_dim;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = lane + 32 * j;
      v[j].x = (v[j].x - m0) * rstd; v[j].y = (v[j].y - m0) * rstd;
      v[j].z = (v[j].z - m0) * rstd; v[j].w = (v[j].w - m0) * rstd;
      acc_g[j].x += d[j].x * v[j].x; acc_g[j].y += d[j].y * v[j].y;
      acc_g[j].z += d[j].z * v[j].z; acc_g[j].w += d[j].w * v[j].w;
      acc_b[j].x += d[j].x; acc_b[j].y += d[j].y; acc_b[j].z += d[j].z; acc_b[j].w += d[j].w;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      float4 o;
      o.x = rstd * (d[j].x * gm.x - mg - v[j].x * mgx);
      o.y = rstd * (d[j].y * gm.y - mg - v[j].y * mgx);
      o.z = rstd * (d[j].z * gm.z - mg - v[j].z * mgx);
      o.w = rstd * (d[j].w * gm.w - mg - v[j].w * mgx);
      if (add_in != nullptr) {
        const float4 a = *reinterpret_cast<const float4*>(b + 2048 + i * 16);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      if (dx_out != nullptr) reinterpret_cast<float4*>(dx_out + row * DIM)[i] = o;
      if (dx_bf16 != nullptr) {
        uint2 pk;
        pk.x = ptx::pack_bf16(o.x, o.y);
        pk.y = ptx::pack_bf16(o.z, o.w);
        reinterpret_cast<uint2*>(dx_bf16 + row * DIM)[i] = pk;
      }
    }
    __syncwarp();            // every lane is done with this stage before the next iteration's cp.async refills it
  }
  cp_async_wait<0>();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    atomicAdd(&red[4 * i + 0], acc_g[j].x); atomicAdd(&red[4 * i + 1], acc_g[j].y);
    atomicAdd(&red[4 * i + 2], acc_g[j].z); atomicAdd(&red[4 * i + 3], acc_g[j].w);
    atomicAdd(&red[DIM + 4 * i + 0], acc_b[j].x); atomicAdd(&red[DIM + 4 * i + 1], acc_b[j].y);
    atomicAdd(&red[DIM + 4 * i + 2], acc_b[j].z); atomicAdd(&red[DIM + 4 * i + 3], acc_b[j].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < DIM; i += blockDim.x) {
    if (dgamma != nullptr) atomicAdd(dgamma + i, red[i]);
    if (dbeta != nullptr) atomicAdd(dbeta + i, red[DIM + i]);
  }
}

// ------------------------------------------------------------------------------------------ GEGLU
// h: [rows][2*ld_half] bf16 = [x | gate];  u: [rows][ld_half] bf16 = x * gelu(gate)
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(const uint4* __restrict__ h, uint4* __restrict__ u, long long rows, int half_vec /* ld_half/8 */) {
  // 32-bit index arithmetic (the host checks rows * half_vec < 2^31): a 64-bit division per vector costs more than the GELUs
  const unsigned total = (unsigned)(rows * half_vec), hv = (unsigned)half_vec;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned row = idx / hv;
    const unsigned c = idx - row * hv;
    const uint4* hr = h + (size_t)row * 2 * hv + c;
    const uint4 xv = hr[0];
    const uint4 gv = hr[hv];
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = ptx::pack_bf16(ptx::bf16_lo(xs[k]) * ptx::gelu_fast(ptx::bf16_lo(gs[k])),
                            ptx::bf16_hi(xs[k]) * ptx::gelu_fast(ptx::bf16_hi(gs[k])));
    u[idx] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// dh = [du * gelu(gate) | du * x * gelu'(gate)]
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const uint4* __restrict__ h, const uint4* __restrict__ du, uint4* __restrict__ dh, long long rows,
                 int half_vec) {
  const unsigned total = (unsigned)(rows * half_vec), hv = (unsigned)half_vec;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned row = idx / hv;
    const unsigned c = idx - row * hv;
    const uint4* hr = h + (size_t)row * 2 * hv + c;
    const uint4 xv = hr[0];
    const uint4 gv = hr[hv];
    const uint4 dv = du[idx];
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w},
                   ds[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t ox[4], og[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x0 = ptx::bf16_lo(xs[k]), x1 = ptx::bf16_hi(xs[k]);
      const float g0 = ptx::bf16_lo(gs[k]), g1 = ptx::bf16_hi(gs[k]);
      const float d0 = ptx::bf16_lo(ds[k]), d1 = ptx::bf16_hi(ds[k]);
      float c0, p0, c1, p1;
      ptx::gelu_parts(g0, c0, p0);
      ptx::gelu_parts(g1, c1, p1);
      ox[k] = ptx::pack_bf16(d0 * (g0 * c0), d1 * (g1 * c1));
      og[k] = ptx::pack_bf16(d0 * x0 * (c0 + p0), d1 * x1 * (c1 + p1));
    }
    uint4* dr = dh + (size_t)row * 2 * hv + c;
    dr[0] = make_uint4(ox[0], ox[1], ox[2], ox[3]);
    dr[hv] = make_uint4(og[0], og[1], og[2], og[3]);
  }
}

// ------------------------------------------------------------------------------------------ casts
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    y[i] = make_uint2(ptx::pack_bf16(v.x, v.y), ptx::pack_bf16(v.z, v.w));
  }
}

int grid_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)ctclip::sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int ctclip_layernorm_fwd(const float* x, long long rows, int dim, const float* gamma, const float* beta,
                                    float eps, void* y_bf16, void* raw_bf16, float* y_f32, void* stream) {
  if (rows <= 0) return CTCLIP_OK;
  if (dim % 4 || dim > kMaxVecAll * 128 || dim <= 0)
    return ctclip::fail(CTCLIP_E_SHAPE, "layernorm_fwd: dim must be a multiple of 4 and <= %d", kMaxVecAll * 128);
  if (x == nullptr || gamma == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "layernorm_fwd: null pointer");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  const int nv = (dim + 127) / 128;
  const int g = grid_for(rows, 8);
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16 *yb = (__nv_bfloat16*)y_bf16, *rb = (__nv_bfloat16*)raw_bf16;
  if (nv <= 1) layernorm_fwd_kernel<1><<<g, 256, 0, s>>>(x, rows, dim, gamma, beta, eps, yb, rb, y_f32);
  else if (nv <= 2) layernorm_fwd_kernel<2><<<g, 256, 0, s>>>(x, rows, dim, gamma, beta, eps, yb, rb, y_f32);
  else if (nv <= 4) layernorm_fwd_kernel<4><<<g, 256, 0, s>>>(x, rows, dim, gamma, beta, eps, yb, rb, y_f32);
  else layernorm_fwd_kernel<8><<<g, 256, 0, s>>>(x, rows, dim, gamma, beta, eps, yb, rb, y_f32);
  return ctclip::check_launch("layernorm_fwd");
}

extern "C" int ctclip_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, long long rows, int dim,
                                    const float* gamma, float eps, const float* add_in, float* dx_out, void* dx_bf16,
                                    float* dgamma, float* dbeta, void* stream) {
  if (rows <= 0) return CTCLIP_OK;
  if (dim % 4 || dim > kMaxVecAll * 128 || dim <= 0)
    return ctclip::fail(CTCLIP_E_SHAPE, "layernorm_bwd: dim must be a multiple of 4 and <= %d", kMaxVecAll * 128);
  if (dy == nullptr || x == nullptr || gamma == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "layernorm_bwd: null pointer");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  // each warp walks >= 8 rows so the column reductions amortise — unless that leaves SMs empty: the BERT tower's [4096, 768]
  // calls got 64 CTAs for 148 SMs (22 us per call, 25 calls per step); two rows per warp fill the chip
  long long blocks = (rows + 63) / 64;
  if (blocks < (long long)ctclip::sm_count() * 2) blocks = (rows + 15) / 16;
  // one resident wave (3 CTAs / SM at <= 80 registers for dim <= 512): the column-sum flush happens once per CTA
  const long long cap = (long long)ctclip::sm_count() * 3;
  if (blocks > cap) blocks = cap;
  const int nv = (dim + 127) / 128;
  const size_t sm = 2 * dim * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16* db = (__nv_bfloat16*)dx_bf16;
  {
    // dim 512, 16-byte aligned rows: the cp.async ring kernel (CTCLIP_LN_BWD_RING=0: the register-staged kernel, A/B)
    const char* e = getenv("CTCLIP_LN_BWD_RING");    // read per call: the test compares both kernels in one process
    const int ring = (e != nullptr && e[0] == '0') ? 0 : 1;
    const uintptr_t al = reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(add_in) |
                         reinterpret_cast<uintptr_t>(dx_out) | reinterpret_cast<uintptr_t>(dx_bf16);
    if (ring && dim == 512 && (al & 15) == 0) {
      long long nb = (rows + 63) / 64;
      const long long cap2 = (long long)ctclip::sm_count() * 2;
      if (nb > cap2) nb = cap2;
      cudaStream_t s2 = (cudaStream_t)stream;
      __nv_bfloat16* db2 = (__nv_bfloat16*)dx_bf16;
      if (dy_is_bf16) {
        constexpr int kSm = 2 * 512 * 4 + 8 * 2 * (4096 + 1024);
        static bool cfg = false;
        if (!cfg) { cudaFuncSetAttribute(layernorm_bwd512_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm); cfg = true; }
        layernorm_bwd512_kernel<true><<<(int)nb, 256, kSm, s2>>>(dy, x, rows, gamma, eps, add_in, dx_out, db2, dgamma, dbeta);
      } else {
        constexpr int kSm = 2 * 512 * 4 + 8 * 2 * (4096 + 2048);
        static bool cfg = false;
        if (!cfg) { cudaFuncSetAttribute(layernorm_bwd512_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm); cfg = true; }
        layernorm_bwd512_kernel<false><<<(int)nb, 256, kSm, s2>>>(dy, x, rows, gamma, eps, add_in, dx_out, db2, dgamma, dbeta);
      }
      return ctclip::check_launch("layernorm_bwd(512)");
    }
  }
  static int ahead = -1;   // CTCLIP_LN_BWD_L2_AHEAD=0: no L2 prefetches (A/B)
  if (ahead < 0) {
    const char* e = getenv("CTCLIP_LN_BWD_L2_AHEAD");
    ahead = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
#define LN_BWD(V, MB)                                                                                                      \
  do {                                                                                                                 \
    if (dy_is_bf16)                                                                                                    \
      layernorm_bwd_kernel<V, true, MB><<<(int)blocks, 256, sm, s>>>(dy, x, rows, dim, gamma, eps, add_in, dx_out, db, dgamma, dbeta, ahead);  \
    else                                                                                                               \
      layernorm_bwd_kernel<V, false, MB><<<(int)blocks, 256, sm, s>>>(dy, x, rows, dim, gamma, eps, add_in, dx_out, db, dgamma, dbeta, ahead); \
  } while (0)
  if (nv <= 1) LN_BWD(1, 3);
  else if (nv <= 2) LN_BWD(2, 3);
  else if (nv <= 4) LN_BWD(4, 3);
  else LN_BWD(8, 1);
#undef LN_BWD
  return ctclip::check_launch("layernorm_bwd");
}

extern "C" int ctclip_geglu_fwd(const void* h, void* u, long long rows, int ld_half, void* stream) {
  if (rows <= 0) return CTCLIP_OK;
  if (ld_half % 8 || ld_half <= 0) return ctclip::fail(CTCLIP_E_ALIGN, "geglu_fwd: ld_half must be a multiple of 8");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  if (rows * (ld_half / 8) >= (1LL << 31)) return ctclip::fail(CTCLIP_E_SHAPE, "geglu_fwd: more than 2^31 vectors");
  geglu_fwd_kernel<<<grid_for(rows * (ld_half / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)h, (uint4*)u, rows, ld_half / 8);
  return ctclip::check_launch("geglu_fwd");
}

extern "C" int ctclip_geglu_bwd(const void* h, const void* du, void* dh, long long rows, int ld_half, void* stream) {
  if (rows <= 0) return CTCLIP_OK;
  if (ld_half % 8 || ld_half <= 0) return ctclip::fail(CTCLIP_E_ALIGN, "geglu_bwd: ld_half must be a multiple of 8");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  if (rows * (ld_half / 8) >= (1LL << 31)) return ctclip::fail(CTCLIP_E_SHAPE, "geglu_bwd: more than 2^31 vectors");
  geglu_bwd_kernel<<<grid_for(rows * (ld_half / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)h, (const uint4*)du, (uint4*)dh, rows, ld_half / 8);
  return ctclip::check_launch("geglu_bwd");
}

extern "C" int ctclip_cast_f32_bf16(const float* x, void* y, long long n, void* stream) {
  if (n <= 0) return CTCLIP_OK;
  if (n % 4) return ctclip::fail(CTCLIP_E_ALIGN, "cast_f32_bf16: n must be a multiple of 4");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  cast_f32_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (uint2*)y, n / 4);
  return ctclip::check_launch("cast_f32_bf16");
}

namespace {
// out[c] += sum_rows x[row][c]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long rows, int dim, float* __restrict__ out, int rows_per_block) {
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += x[r * dim + c];
    atomicAdd(out + c, s);
  }
}
// batched 2-D bf16 copies (operand re-packing: zero-padded FF weights): desc[i] = {src, dst, rows, cols, src_ld, dst_ld}
__global__ void __launch_bounds__(256)
copy2d_batch_kernel(const long long* __restrict__ desc) {
  const long long* d = desc + 6 * blockIdx.y;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(d[0]);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(d[1]);
  const long long rows = d[2], cols = d[3], sld = d[4], dld = d[5];
  const bool vec = (cols % 8 == 0) && (sld % 8 == 0) && (dld % 8 == 0) && ((d[0] | d[1]) % 16 == 0);
  if (vec) {
    const long long cv = cols / 8, total = rows * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / cv, c = i - r * cv;
      *reinterpret_cast<uint4*>(dst + r * dld + 8 * c) = *reinterpret_cast<const uint4*>(src + r * sld + 8 * c);
    }
  } else {
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / cols, c = i - r * cols;
      dst[r * dld + c] = src[r * sld + c];
    }
  }
}

__global__ void __launch_bounds__(256)
colsum_tile_kernel(const float* __restrict__ x, long long rows, int dim, float* __restrict__ out, int rows_per_block) {
  ptx::colsum_tile<float>(x, rows, dim, dim, out, rows_per_block);
}
}  // namespace

extern "C" int ctclip_colsum(const float* x, long long rows, int dim, float* out, void* stream) {
  if (rows <= 0 || dim <= 0) return CTCLIP_OK;
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  if (dim % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int bx = (dim / 4 + 255) / 256;
    long long rpb8 = rows * bx / (2LL * ctclip::sm_count());   // about two CTAs per SM
    rpb8 = rpb8 < 8 ? 8 : (rpb8 > 64 ? 64 : rpb8 / 8 * 8);
    dim3 grid((unsigned)bx, (unsigned)((rows + rpb8 - 1) / rpb8));
    colsum_tile_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, dim, out, (int)rpb8);
    return ctclip::check_launch("colsum");
  }
  int rpb = 64;
  long long blocks = (rows + rpb - 1) / rpb;
  const long long cap = (long long)ctclip::sm_count() * 8;
  if (blocks > cap) {
    rpb = (int)((rows + cap - 1) / cap);
    blocks = (rows + rpb - 1) / rpb;
  }
  colsum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, dim, out, rpb);
  return ctclip::check_launch("colsum");
}

extern "C" int ctclip_copy2d_batch_bf16(const long long* desc_dev, int n, void* stream) {
  if (n <= 0) return CTCLIP_OK;
  if (desc_dev == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "copy2d_batch: null descriptor table");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  copy2d_batch_kernel<<<dim3((unsigned)ctclip::sm_count(), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(desc_dev);
  return ctclip::check_launch("copy2d_batch");
}
