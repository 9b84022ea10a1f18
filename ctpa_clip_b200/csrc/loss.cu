// Symmetric InfoNCE over the (global) batch of l2-normalised latents (reference ct_clip.py:771,796,845-878):
//   L = exp(tau) T I^T ; loss = mean_i( -L_ii + lse_j L_ij )/2 + mean_j( -L_jj + lse_i L_ij )/2
// plus its gradient w.r.t. the latents of this rank's rows [row0, row0+rows_local) and this rank's share of dtau.
// fp32 on CUDA cores: B <= a few hundred, 2*B^2*512 flop is noise next to the encoders; warp-shuffle reductions.
// In data-parallel runs T / I are the all-gathered latents, so every rank forms the global-batch logits.
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// L[i][j] = exp(tau) * <T_i, I_j>; one warp per (i, j)
__global__ void __launch_bounds__(256)
clip_logits_kernel(const float* __restrict__ T, const float* __restrict__ I, const float* __restrict__ tau, int B, int d,
                   float* __restrict__ L) {
  const int lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= (long long)B * B) return;
  const int i = (int)(pair / B), j = (int)(pair % B);
  float s = 0.f;
  for (int k = lane * 4; k < d; k += 128) {
    const float4 a = *reinterpret_cast<const float4*>(T + (long long)i * d + k);
    const float4 b = *reinterpret_cast<const float4*>(I + (long long)j * d + k);
    s += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
  }
  s = warp_sum(s);
  if (lane == 0) L[pair] = __expf(*tau) * s;
}

// lse[0][i] = logsumexp_j L[i][j] (text->image rows); lse[1][j] = logsumexp_i L[i][j] (image->text)
__global__ void __launch_bounds__(256)
clip_lse_kernel(const float* __restrict__ L, int B, float* __restrict__ lse) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= 2 * B) return;
  const bool col = w >= B;
  const int r = col ? w - B : w;
  float m = -INFINITY;
  for (int k = lane; k < B; k += 32) m = fmaxf(m, col ? L[(long long)k * B + r] : L[(long long)r * B + k]);
  m = warp_max(m);
  float s = 0.f;
  for (int k = lane; k < B; k += 32) s += __expf((col ? L[(long long)k * B + r] : L[(long long)r * B + k]) - m);
  s = warp_sum(s);
  if (lane == 0) lse[w] = m + __logf(s);
}

// block b < rows_local: dT[b] ; block b >= rows_local: dI[b - rows_local]. Block 0 also writes the loss.
// G_ij = (exp(L_ij - lse_row_i) + exp(L_ij - lse_col_j) - 2 delta_ij) / (2B)
__global__ void __launch_bounds__(128)
clip_grad_kernel(const float* __restrict__ T, const float* __restrict__ I, const float* __restrict__ L,
                 const float* __restrict__ lse, const float* __restrict__ tau, int B, int d, int row0, int rows_local,
                 float* __restrict__ dT, float* __restrict__ dI, float* __restrict__ loss, float* __restrict__ dtau) {
  extern __shared__ float coef[];  // [B]
  __shared__ float red[4];
  const bool is_img = blockIdx.x >= rows_local;
  const int r = row0 + (is_img ? blockIdx.x - rows_local : blockIdx.x);
  const float inv2b = 0.5f / B;
  const float et = __expf(*tau);
  float part = 0.f;
  for (int k = threadIdx.x; k < B; k += blockDim.x) {
    const int i = is_img ? k : r, j = is_img ? r : k;
    const float l = L[(long long)i * B + j];
    const float g = (__expf(l - lse[i]) + __expf(l - lse[B + j]) - (i == j ? 2.f : 0.f)) * inv2b;
    coef[k] = g * et;
    if (!is_img) part += g * l;  // dtau share of text row r
  }
  __syncthreads();
  const float* src = is_img ? T : I;
  float* dst = (is_img ? dI : dT) + (long long)(r - row0) * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < B; ++k) acc = fmaf(coef[k], src[(long long)k * d + c], acc);
    dst[c] = acc;
  }
  if (!is_img && dtau != nullptr) {
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(dtau, red[0] + red[1] + red[2] + red[3]);
  }
  if (blockIdx.x == 0 && loss != nullptr) {
    __syncthreads();
    float s = 0.f;
    for (int k = threadIdx.x; k < B; k += blockDim.x) s += lse[k] + lse[B + k] - 2.f * L[(long long)k * B + k];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) *loss = (red[0] + red[1] + red[2] + red[3]) * inv2b;
  }
}

// zero-shot scoring (ctclip_inference.py:286-336): for volume v and pathology p the reference forms the two logits
// exp(tau) <I_v, T_{2p}> ("present") and exp(tau) <I_v, T_{2p+1}> ("not present") and keeps softmax(.)[0].
// One warp per (v, p).
__global__ void __launch_bounds__(256)
zero_shot_kernel(const float* __restrict__ I, const float* __restrict__ T, const float* __restrict__ tau, int V, int P, int d,
                 float* __restrict__ prob, float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= (long long)V * P) return;
  const int v = (int)(item / P), p = (int)(item % P);
  const float* iv = I + (long long)v * d;
  const float* t0 = T + (long long)(2 * p) * d;
  const float* t1 = t0 + d;
  float s0 = 0.f, s1 = 0.f;
  for (int k = lane * 4; k < d; k += 128) {
    const float4 a = *reinterpret_cast<const float4*>(iv + k);
    const float4 b = *reinterpret_cast<const float4*>(t0 + k);
    const float4 c = *reinterpret_cast<const float4*>(t1 + k);
    s0 += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
    s1 += (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w);
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) {
    const float et = __expf(*tau);
    const float l0 = et * s0, l1 = et * s1;
    const float m = fmaxf(l0, l1);
    const float e0 = __expf(l0 - m), e1 = __expf(l1 - m);
    prob[item] = e0 / (e0 + e1);
    if (logits != nullptr) {
      logits[2 * item] = l0;
      logits[2 * item + 1] = l1;
    }
  }
}

// y = x / max(|x|, eps):  dx = (g - y <y, g>) * inv_norm          (F.normalize backward)
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ inv_norm, const float* __restrict__ g,
                  long long rows, int dim, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float dot = 0.f;
  for (int k = lane; k < dim; k += 32) dot = fmaf(y[row * dim + k], g[row * dim + k], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  for (int k = lane; k < dim; k += 32) dx[row * dim + k] = (g[row * dim + k] - y[row * dim + k] * dot) * inv;
}

}  // namespace

namespace ctclip {
// logits already in work[0, B*B): row/column log-sum-exp, then loss + gradient of rows [row0, row0+rows_local).
// Shared by ctclip_clip_loss and the peer-memory path (symm.cu: ctclip_clip_loss_allgather).
int clip_lse_grad_launch(const float* T, const float* I, const float* tau, int B, int d, int row0, int rows_local,
                         float* work, float* loss, float* dT, float* dI, float* dtau, cudaStream_t s) {
  float* L = work;
  float* lse = work + (long long)B * B;
  clip_lse_kernel<<<(2 * B + 7) / 8, 256, 0, s>>>(L, B, lse);
  int rc = ctclip::check_launch("clip_lse");
  if (rc) return rc;
  const int blocks = (dT != nullptr && dI != nullptr) ? 2 * rows_local : 0;
  if (blocks > 0) {
    clip_grad_kernel<<<blocks, 128, B * sizeof(float), s>>>(T, I, L, lse, tau, B, d, row0, rows_local, dT, dI, loss, dtau);
  } else {
    // loss only: a single "text row" block whose gradient rows land in the scratch tail of `work`
    clip_grad_kernel<<<1, 128, B * sizeof(float), s>>>(T, I, L, lse, tau, B, d, 0, 1, work + (long long)B * B + 2 * B,
                                                       work + (long long)B * B + 2 * B, loss, nullptr);
  }
  return ctclip::check_launch("clip_grad");
}
}  // namespace ctclip

// T, I: fp32 [B][d] l2-normalised latents of the whole (global) batch; tau: device scalar (temperature parameter).
// work: fp32 [B*B + 2*B]. Outputs: loss (device scalar), dT/dI fp32 [rows_local][d] (grad w.r.t. the normalised latents of
// rows [row0, row0+rows_local)), dtau (+= this rank's share). Any of loss/dtau may be NULL.
extern "C" int ctclip_clip_loss(const float* T, const float* I, const float* tau, int B, int d, int row0, int rows_local,
                                float* work, float* loss, float* dT, float* dI, float* dtau, void* stream) {
  if (B <= 0 || d <= 0 || d % 4) return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss: bad shape B=%d d=%d", B, d);
  if (row0 < 0 || rows_local < 0 || row0 + rows_local > B) return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss: bad local row range");
  if (B > 8192) return ctclip::fail(CTCLIP_E_SHAPE, "clip_loss: batch %d too large", B);
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  clip_logits_kernel<<<(unsigned)(((long long)B * B + 7) / 8), 256, 0, s>>>(T, I, tau, B, d, work);
  rc = ctclip::check_launch("clip_logits");
  if (rc) return rc;
  return ctclip::clip_lse_grad_launch(T, I, tau, B, d, row0, rows_local, work, loss, dT, dI, dtau, s);
}

extern "C" int ctclip_l2norm_bwd(const float* y, const float* inv_norm, const float* g, long long rows, int dim, float* dx,
                                 void* stream) {
  if (rows <= 0) return CTCLIP_OK;
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  l2norm_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(y, inv_norm, g, rows, dim, dx);
  return ctclip::check_launch("l2norm_bwd");
}

// I: fp32 [V][d] normalised image latents; T: fp32 [2P][d] normalised text latents ordered [present_0, absent_0, ...];
// prob: fp32 [V][P] = softmax pair [0]; logits (may be NULL): fp32 [V][P][2].
extern "C" int ctclip_zero_shot_scores(const float* I, const float* T, const float* tau, int V, int P, int d, float* prob,
                                       float* logits, void* stream) {
  if (V <= 0 || P <= 0 || d <= 0 || d % 4) return ctclip::fail(CTCLIP_E_SHAPE, "zero_shot_scores: bad shape V=%d P=%d d=%d", V, P, d);
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  zero_shot_kernel<<<(unsigned)(((long long)V * P + 7) / 8), 256, 0, (cudaStream_t)stream>>>(I, T, tau, V, P, d, prob, logits);
  return ctclip::check_launch("zero_shot_scores");
}
