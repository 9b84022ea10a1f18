// Optimiser step of the reference trainer on flat fp32 arenas (reference CTCLIPTrainer.py:347-353, optimizer.py:10-24):
// global-norm gradient clipping (max_norm 0.5) + torch.optim.Adam (betas 0.9/0.99, eps 1e-8, no weight decay), fused into
// one HBM pass that also refreshes the bf16 shadow copy used as tensor-core operand and zeroes the gradient.
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {

constexpr int kSumsqBlocks = 1024;   // fixed grid: the partial sums and their order do not depend on the device or the launch

// Deterministic sum of squares: block b writes ONE partial (fixed strided element order, shuffle tree, fixed warp order) and the
// second kernel adds the kSumsqBlocks partials in index order. No atomics: every data-parallel rank gets the bit-identical
// norm from bit-identical (all-reduced) gradients, so the clip factor — and with it the replicas — cannot drift apart.
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float4* __restrict__ g, long long nvec, float* __restrict__ partials) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  __shared__ float sh[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] += sh[0];
}

__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
            uint2* __restrict__ shadow, long long nvec, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
            const float* __restrict__ norm_sq, float max_norm, int zero_grad, int* __restrict__ skipped) {
  float clip = 1.f;
  if (norm_sq != nullptr) {
    const float nsq = *norm_sq;
    // NaN / Inf gradients (a peer that timed out in the latent exchange, an overflow): never a training update. Nothing is
    // written — parameters, moments, bf16 mirror and the gradients themselves stay for the host to inspect.
    if (!(fabsf(nsq) <= 3.0e38f)) {
      if (skipped != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(skipped, 1);
      return;
    }
    if (max_norm > 0.f) {
      const float c = max_norm / (sqrtf(nsq) + 1e-6f);  // torch.nn.utils.clip_grad_norm_
      clip = c < 1.f ? c : 1.f;
    }
  }
  const float step = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G[k] * clip;
      M[k] = b1 * M[k] + (1.f - b1) * gr;
      V[k] = b2 * V[k] + (1.f - b2) * gr * gr;
      P[k] -= step * M[k] / (sqrtf(V[k]) / bc2_sqrt + eps);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow != nullptr) shadow[i] = make_uint2(ptx::pack_bf16(pp.x, pp.y), ptx::pack_bf16(pp.z, pp.w));
  }
}

}  // namespace

// out[0] += sum(g^2), deterministically (see above). workspace: device scratch of ctclip_workspace_bytes("sumsq") bytes.
extern "C" int ctclip_sumsq(const float* g, long long n, float* out, float* workspace, void* stream) {
  if (n <= 0) return CTCLIP_OK;
  if (n % 4 || (reinterpret_cast<uintptr_t>(g) & 15)) return ctclip::fail(CTCLIP_E_ALIGN, "sumsq: n %% 4 and 16-byte alignment required");
  if (workspace == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "sumsq: workspace of ctclip_workspace_bytes(\"sumsq\") bytes required");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  sumsq_partial_kernel<<<kSumsqBlocks, 256, 0, (cudaStream_t)stream>>>((const float4*)g, n / 4, workspace);
  rc = ctclip::check_launch("sumsq(partials)");
  if (rc) return rc;
  sumsq_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(workspace, kSumsqBlocks, out);
  return ctclip::check_launch("sumsq(final)");
}

// one Adam step over flat arenas; `step` is the 1-based step count; norm_sq (device scalar, may be NULL) enables clipping
extern "C" int ctclip_adam_step(float* p, float* g, float* m, float* v, void* bf16_shadow, long long n, float lr, float beta1,
                                float beta2, float eps, int step, const float* norm_sq, float max_norm, int zero_grad,
                                int* skipped, void* stream) {
  if (n <= 0) return CTCLIP_OK;
  if (n % 4) return ctclip::fail(CTCLIP_E_ALIGN, "adam_step: n must be a multiple of 4");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)ctclip::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v,
                                                                 (uint2*)bf16_shadow, n / 4, lr, beta1, beta2, eps, bc1,
                                                                 bc2_sqrt, norm_sq, max_norm, zero_grad, skipped);
  return ctclip::check_launch("adam_step");
}
