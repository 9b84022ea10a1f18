// 3-D tubelet patchify fused with the first LayerNorm of CTViT.to_patch_emb (reference ctvit.py:170-171):
//   'b c (t pt) (h p1) (w p2) -> b t h w (c pt p1 p2)'  ->  LayerNorm(pt*p1*p2, affine, eps 1e-5)  -> bf16 GEMM operand
// One CTA per token: the patch (pt x p1 x p2 fp32, p2 contiguous in HBM) is gathered once into registers,
// mean / variance are block-reduced, and the normalised row is written as one contiguous bf16 line (the A operand of
// the patch-embed tcgen05 GEMM).
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxPer = 16;  // patch elements per thread -> patch dim <= 4096

struct PatchGeom {
  int frames, height, width;  // volume (c = 1)
  int pt, ps;                 // temporal / spatial patch size
  int gt, gh, gw;             // token grid
  int pdim;                   // pt * ps * ps
};

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ const float* patch_base(const float* video, const PatchGeom& g, long long token) {
  const int per_b = g.gt * g.gh * g.gw;
  const long long b = token / per_b;
  int c = (int)(token - b * per_b);
  const int wi = c % g.gw; c /= g.gw;
  const int hi = c % g.gh;
  const int ti = c / g.gh;
  return video + ((b * g.frames + (long long)ti * g.pt) * g.height + (long long)hi * g.ps) * g.width + (long long)wi * g.ps;
}
// element e = (dt*ps + dy)*ps + dx of the patch
__device__ __forceinline__ long long patch_elem_off(const PatchGeom& g, int e) {
  const int dx = e % g.ps;
  const int r = e / g.ps;
  const int dy = r % g.ps, dt = r / g.ps;
  return ((long long)dt * g.height + dy) * g.width + dx;
}

__global__ void __launch_bounds__(kThreads)
patch_ln_fwd_kernel(const float* __restrict__ video, __nv_bfloat16* __restrict__ out, long long ld_out,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, PatchGeom g) {
  __shared__ float red[kThreads / 32];
  const long long token = blockIdx.x;
  const float* base = patch_base(video, g, token);
  float v[kMaxPer];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int e = threadIdx.x + j * kThreads;
    v[j] = (e < g.pdim) ? base[patch_elem_off(g, e)] : 0.f;
    s += v[j];
  }
  const float mean = block_sum(s, red) / g.pdim;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int e = threadIdx.x + j * kThreads;
    const float d = (e < g.pdim) ? v[j] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(block_sum(q, red) / g.pdim + eps);
  __nv_bfloat16* orow = out + token * ld_out;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int e = threadIdx.x + j * kThreads;
    if (e < g.pdim) orow[e] = __float2bfloat16_rn((v[j] - mean) * rstd * gamma[e] + beta[e]);
  }
}

// 128-bit variant (patch width % 4 == 0, so every 4-element group lies inside one contiguous w-run of the volume)
__global__ void __launch_bounds__(kThreads)
patch_ln_fwd_vec4_kernel(const float* __restrict__ video, __nv_bfloat16* __restrict__ out, long long ld_out,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps, PatchGeom g) {
  __shared__ float red[kThreads / 32];
  const long long token = blockIdx.x;
  const float* base = patch_base(video, g, token);
  constexpr int kSlots = kMaxPer / 4;
  const int nvec = g.pdim >> 2;
  float4 v[kSlots];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kSlots; ++j) {
    const int f = threadIdx.x + j * kThreads;
    v[j] = (f < nvec) ? *reinterpret_cast<const float4*>(base + patch_elem_off(g, 4 * f)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = block_sum(s, red) / g.pdim;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kSlots; ++j) {
    const int f = threadIdx.x + j * kThreads;
    if (f < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(block_sum(q, red) / g.pdim + eps);
  __nv_bfloat16* orow = out + token * ld_out;
#pragma unroll
  for (int j = 0; j < kSlots; ++j) {
    const int f = threadIdx.x + j * kThreads;
    if (f < nvec) {
      const float4 gm = reinterpret_cast<const float4*>(gamma)[f];
      const float4 bt = reinterpret_cast<const float4*>(beta)[f];
      const float o0 = (v[j].x - mean) * rstd * gm.x + bt.x, o1 = (v[j].y - mean) * rstd * gm.y + bt.y;
      const float o2 = (v[j].z - mean) * rstd * gm.z + bt.z, o3 = (v[j].w - mean) * rstd * gm.w + bt.w;
      reinterpret_cast<uint2*>(orow)[f] = make_uint2(ptx::pack_bf16(o0, o1), ptx::pack_bf16(o2, o3));
    }
  }
}

// LayerNorm(pdim) parameter gradients WITHOUT the tokens x pdim dgrad GEMM. With A = xhat*gamma + beta (the GEMM operand),
// Y = A W^T + b and s = colsum(dY) (= the Linear's bias gradient):
//   dW[n][e]  = sum_tok dY[tok][n] A[tok][e] = gamma[e] * M[n][e] + beta[e] * s[n],   M = dY^T xhat
//   dbeta[e]  = sum_n s[n] W[n][e]
//   dgamma[e] = sum_n W[n][e] M[n][e] = sum_n W[n][e] (dW[n][e] - beta[e] s[n]) / gamma[e]
// so both follow from the weight gradient that the wgrad GEMM produces anyway.
__global__ void __launch_bounds__(256)
patch_ln_param_grad_kernel(const float* __restrict__ W, const float* __restrict__ dW, const float* __restrict__ s,
                           const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dgamma,
                           float* __restrict__ dbeta, int n_out, int pdim) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= pdim) return;
  const float be = beta[e];
  float accg = 0.f, accb = 0.f;
  for (int n = 0; n < n_out; ++n) {
    const float w = W[(long long)n * pdim + e];
    const float sn = s[n];
    accb = fmaf(sn, w, accb);
    accg = fmaf(w, dW[(long long)n * pdim + e] - be * sn, accg);
  }
  dgamma[e] += accg / gamma[e];
  dbeta[e] += accb;
}

int make_geom(PatchGeom& g, int frames, int height, int width, int pt, int ps, const char* what) {
  if (pt <= 0 || ps <= 0 || frames % pt || height % ps || width % ps)
    return ctclip::fail(CTCLIP_E_SHAPE, "%s: volume %dx%dx%d is not divisible by the patch %dx%dx%d", what, frames,
                        height, width, pt, ps, ps);
  g.frames = frames; g.height = height; g.width = width; g.pt = pt; g.ps = ps;
  g.gt = frames / pt; g.gh = height / ps; g.gw = width / ps;
  g.pdim = pt * ps * ps;
  if (g.pdim > kMaxPer * kThreads) return ctclip::fail(CTCLIP_E_SHAPE, "%s: patch dim %d > %d", what, g.pdim, kMaxPer * kThreads);
  return ctclip::require_sm100();
}

}  // namespace

// video fp32 [batch][1][frames][height][width] -> out bf16 [batch*gt*gh*gw][ld_out] (first pdim columns written)
extern "C" int ctclip_patch_ln_fwd(const float* video, int batch, int frames, int height, int width, int pt, int ps,
                                   const float* gamma, const float* beta, float eps, void* out, long long ld_out,
                                   void* stream) {
  PatchGeom g;
  int rc = make_geom(g, frames, height, width, pt, ps, "patch_ln_fwd");
  if (rc) return rc;
  if (batch <= 0) return CTCLIP_OK;
  const long long tokens = (long long)batch * g.gt * g.gh * g.gw;
  const bool vec4 = (ps % 4 == 0) && (width % 4 == 0) && (ld_out % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(video) | reinterpret_cast<uintptr_t>(gamma) |
                      reinterpret_cast<uintptr_t>(beta)) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 8 == 0);
  if (vec4)
    patch_ln_fwd_vec4_kernel<<<(unsigned)tokens, kThreads, 0, (cudaStream_t)stream>>>(video, (__nv_bfloat16*)out, ld_out,
                                                                                     gamma, beta, eps, g);
  else
    patch_ln_fwd_kernel<<<(unsigned)tokens, kThreads, 0, (cudaStream_t)stream>>>(video, (__nv_bfloat16*)out, ld_out, gamma,
                                                                                beta, eps, g);
  return ctclip::check_launch("patch_ln_fwd");
}

// dgamma/dbeta (+=) of the patch LayerNorm from the patch-embed Linear's W, dW and bias gradient s (all fp32)
extern "C" int ctclip_patch_ln_param_grad(const float* W, const float* dW, const float* s, const float* gamma,
                                          const float* beta, float* dgamma, float* dbeta, int n_out, int pdim,
                                          void* stream) {
  if (n_out <= 0 || pdim <= 0) return ctclip::fail(CTCLIP_E_SHAPE, "patch_ln_param_grad: empty problem");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  patch_ln_param_grad_kernel<<<(pdim + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, dW, s, gamma, beta, dgamma, dbeta,
                                                                                  n_out, pdim);
  return ctclip::check_launch("patch_ln_param_grad");
}
