// Cosine-sim vector quantiser of CTViT (vector_quantize_pytorch 1.1.2 semantics; call site ctvit.py:427):
//   idx = argmax_c <l2norm(x), l2norm(e_c)> ; quantize = e_idx ; train: EMA update of (cluster_size, embed).
// The T x C score matrix is never materialised: the tcgen05 GEMM (top-2 epilogue, gemm_sm100.cu) emits per
// (token, 256-code tile) the two best *bf16-operand* scores; vq_finalize then re-scores in exact fp32 every code whose
// coarse score is within `margin` of the coarse maximum (margin >= 2x the provable bf16 rounding bound 2^-8 of a unit
// dot product), so the arg-max equals the fp32 arg-max regardless of tensor-core rounding.
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

constexpr int kMaxVec = 8;  // dim <= 1024

// y = x / max(|x|, 1e-12) per row; optional fp32 / bf16 outputs and the inverse norm
__global__ void __launch_bounds__(256)
l2norm_rows_kernel(const float* __restrict__ x, long long rows, int dim, float* __restrict__ y_f32,
                   __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = dim >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
  float4 v[kMaxVec];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      v[j] = xr[i];
      ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  if (lane == 0 && inv_norm != nullptr) inv_norm[row] = inv;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      const float4 o = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
      if (y_f32 != nullptr) reinterpret_cast<float4*>(y_f32 + row * dim)[i] = o;
      if (y_bf16 != nullptr)
        reinterpret_cast<uint2*>(y_bf16 + row * dim)[i] = make_uint2(ptx::pack_bf16(o.x, o.y), ptx::pack_bf16(o.z, o.w));
    }
  }
}

// one warp per token: exact fp32 re-score of the near-maximal coarse candidates
__global__ void __launch_bounds__(256)
vq_finalize_kernel(const float4* __restrict__ top2, int n_tiles, int tile_n, const float* __restrict__ x,
                   const float* __restrict__ embed_n, long long rows, int dim, int codes, float margin,
                   int* __restrict__ idx_out, unsigned long long* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = dim >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
  float4 v[kMaxVec];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) v[j] = xr[i];
  }
  float m = -INFINITY;
  for (int t = lane; t < n_tiles; t += 32) m = fmaxf(m, top2[row * n_tiles + t].x);
  m = warp_max(m);
  const float thr = m - margin;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  unsigned long long n_cand = 0, n_scan = 0;
  auto exact = [&](int code) {
    const float4* er = reinterpret_cast<const float4*>(embed_n + (long long)code * dim);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        const float4 e = er[i];
        s += (v[j].x * e.x + v[j].y * e.y) + (v[j].z * e.z + v[j].w * e.w);
      }
    }
    s = warp_sum(s);
    if (s > best || (s == best && code < best_i)) { best = s; best_i = code; }
    ++n_cand;
  };
  for (int t0 = 0; t0 < n_tiles; t0 += 32) {
    const int t = t0 + lane;
    float4 e = make_float4(-INFINITY, 0.f, -INFINITY, 0.f);
    if (t < n_tiles) e = top2[row * n_tiles + t];
    const unsigned has1 = __ballot_sync(0xffffffffu, e.x >= thr);
    const unsigned has2 = __ballot_sync(0xffffffffu, e.z >= thr);
    for (int l = 0; l < 32; ++l) {
      if (!((has1 >> l) & 1u)) continue;
      if ((has2 >> l) & 1u) {
        // two near-maximal codes in one tile: a third may hide behind them -> exact scan of the whole tile
        const int c0 = (t0 + l) * tile_n;
        const int c1 = min(codes, c0 + tile_n);
        for (int code = c0; code < c1; ++code) exact(code);
        ++n_scan;
      } else {
        exact(__float_as_int(__shfl_sync(0xffffffffu, e.y, l)));
      }
    }
  }
  if (lane == 0) {
    idx_out[row] = best_i;
    if (stats != nullptr) {
      atomicAdd(stats + 0, n_cand);
      atomicAdd(stats + 1, n_scan);
    }
  }
}

// out[b][p][:] = (1/t) sum_ti embed[idx[b][ti][p]][:]      (mean over t of the quantised tokens, ct_clip.py:724)
__global__ void __launch_bounds__(128)
vq_gather_mean_kernel(const float* __restrict__ embed, const int* __restrict__ idx, float* __restrict__ out_f32,
                      __nv_bfloat16* __restrict__ out_bf16, int t, int hw, int dim) {
  const long long bp = blockIdx.x;  // b*hw + p
  const long long b = bp / hw;
  const int pp = (int)(bp - b * hw);
  const float inv_t = 1.f / t;
  for (int i = threadIdx.x; i < (dim >> 2); i += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ti = 0; ti < t; ++ti) {
      const int code = idx[(b * t + ti) * hw + pp];
      const float4 e = reinterpret_cast<const float4*>(embed + (long long)code * dim)[i];
      acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
    }
    acc.x *= inv_t; acc.y *= inv_t; acc.z *= inv_t; acc.w *= inv_t;
    if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32 + bp * dim)[i] = acc;
    if (out_bf16 != nullptr)
      reinterpret_cast<uint2*>(out_bf16 + bp * dim)[i] =
          make_uint2(ptx::pack_bf16(acc.x, acc.y), ptx::pack_bf16(acc.z, acc.w));
  }
}

__global__ void __launch_bounds__(128)
vq_gather_kernel(const float* __restrict__ embed, const int* __restrict__ idx, float* __restrict__ out, int dim) {
  const long long row = blockIdx.x;
  const int code = idx[row];
  for (int i = threadIdx.x; i < (dim >> 2); i += blockDim.x)
    reinterpret_cast<float4*>(out + row * dim)[i] = reinterpret_cast<const float4*>(embed + (long long)code * dim)[i];
}

// dx[b][ti][p][:] = dpool[b][p][:] / t     (straight-through estimator + mean-over-t backward)
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* __restrict__ dpool, float* __restrict__ dx, int t, int hw, int dim, long long total_vec) {
  const int dv = dim >> 2;
  const float inv_t = 1.f / t;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv);
    const long long tok = i / dv;
    const int pp = (int)(tok % hw);
    const long long b = tok / ((long long)hw * t);
    float4 g = reinterpret_cast<const float4*>(dpool)[(b * hw + pp) * dv + c];
    g.x *= inv_t; g.y *= inv_t; g.z *= inv_t; g.w *= inv_t;
    reinterpret_cast<float4*>(dx)[i] = g;
  }
}

// EMA statistics: bins[c] += 1, embed_sum[c][:] += l2norm(x_row)
__global__ void __launch_bounds__(256)
vq_ema_accum_kernel(const float* __restrict__ x, const float* __restrict__ inv_norm, const int* __restrict__ idx,
                    long long rows, int dim, float* __restrict__ bins, float* __restrict__ embed_sum) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int code = idx[row];
  const float inv = inv_norm[row];
  if (lane == 0) atomicAdd(bins + code, 1.f);
  for (int i = lane; i < dim; i += 32) atomicAdd(embed_sum + (long long)code * dim + i, x[row * dim + i] * inv);
}

// embed <- lerp(embed, l2norm(embed_sum / bins) (or l2norm(embed) where bins == 0), 1 - decay); cluster_size likewise
__global__ void __launch_bounds__(256)
vq_ema_update_kernel(float* __restrict__ embed, float* __restrict__ cluster_size, const float* __restrict__ bins,
                     const float* __restrict__ embed_sum, int codes, int dim, float decay) {
  const int lane = threadIdx.x & 31;
  const int code = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (code >= codes) return;
  const float n = bins[code];
  const float w = 1.f - decay;
  const bool empty = (n == 0.f);
  const float* src = empty ? (embed + (long long)code * dim) : (embed_sum + (long long)code * dim);
  const float div = empty ? 1.f : n;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) {
    const float a = src[i] / div;
    ss = fmaf(a, a, ss);
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  for (int i = lane; i < dim; i += 32) {
    const float target = (src[i] / div) * inv;
    const float e = embed[(long long)code * dim + i];
    embed[(long long)code * dim + i] = e + w * (target - e);
  }
  if (lane == 0) cluster_size[code] = cluster_size[code] + w * (n - cluster_size[code]);
}

int check_dim(int dim, const char* what) {
  if (dim <= 0 || dim % 4 || dim > kMaxVec * 128) return ctclip::fail(CTCLIP_E_SHAPE, "%s: dim must be a multiple of 4 and <= 1024", what);
  return ctclip::require_sm100();
}

}  // namespace

extern "C" int ctclip_l2norm_rows(const float* x, long long rows, int dim, float* y_f32, void* y_bf16, float* inv_norm,
                                  void* stream) {
  int rc = check_dim(dim, "l2norm_rows");
  if (rc) return rc;
  if (rows <= 0) return CTCLIP_OK;
  l2norm_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, dim, y_f32,
                                                                                   (__nv_bfloat16*)y_bf16, inv_norm);
  return ctclip::check_launch("l2norm_rows");
}

extern "C" int ctclip_vq_finalize(const void* top2, int n_tiles, int tile_n, const float* x, const float* embed_n,
                                  long long rows, int dim, int codes, float margin, int* idx_out,
                                  unsigned long long* stats, void* stream) {
  int rc = check_dim(dim, "vq_finalize");
  if (rc) return rc;
  if (rows <= 0) return CTCLIP_OK;
  vq_finalize_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)top2, n_tiles, tile_n, x, embed_n, rows, dim, codes, margin, idx_out, stats);
  return ctclip::check_launch("vq_finalize");
}

extern "C" int ctclip_vq_gather_mean(const float* embed, const int* idx, float* out_f32, void* out_bf16, int batch, int t,
                                     int hw, int dim, void* stream) {
  int rc = check_dim(dim, "vq_gather_mean");
  if (rc) return rc;
  if (batch <= 0) return CTCLIP_OK;
  vq_gather_mean_kernel<<<(unsigned)((long long)batch * hw), 128, 0, (cudaStream_t)stream>>>(
      embed, idx, out_f32, (__nv_bfloat16*)out_bf16, t, hw, dim);
  return ctclip::check_launch("vq_gather_mean");
}

extern "C" int ctclip_vq_gather(const float* embed, const int* idx, float* out, long long rows, int dim, void* stream) {
  int rc = check_dim(dim, "vq_gather");
  if (rc) return rc;
  if (rows <= 0) return CTCLIP_OK;
  vq_gather_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(embed, idx, out, dim);
  return ctclip::check_launch("vq_gather");
}

extern "C" int ctclip_pool_bwd(const float* dpool, float* dx, int batch, int t, int hw, int dim, void* stream) {
  int rc = check_dim(dim, "pool_bwd");
  if (rc) return rc;
  const long long total_vec = (long long)batch * t * hw * (dim / 4);
  if (total_vec <= 0) return CTCLIP_OK;
  long long blocks = (total_vec + 255) / 256;
  const long long cap = (long long)ctclip::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  pool_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dpool, dx, t, hw, dim, total_vec);
  return ctclip::check_launch("pool_bwd");
}

extern "C" int ctclip_vq_ema_accum(const float* x, const float* inv_norm, const int* idx, long long rows, int dim,
                                   float* bins, float* embed_sum, void* stream) {
  int rc = check_dim(dim, "vq_ema_accum");
  if (rc) return rc;
  if (rows <= 0) return CTCLIP_OK;
  vq_ema_accum_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, inv_norm, idx, rows, dim, bins,
                                                                                    embed_sum);
  return ctclip::check_launch("vq_ema_accum");
}

extern "C" int ctclip_vq_ema_update(float* embed, float* cluster_size, const float* bins, const float* embed_sum,
                                    int codes, int dim, float decay, void* stream) {
  int rc = check_dim(dim, "vq_ema_update");
  if (rc) return rc;
  if (codes <= 0) return CTCLIP_OK;
  vq_ema_update_kernel<<<(codes + 7) / 8, 256, 0, (cudaStream_t)stream>>>(embed, cluster_size, bins, embed_sum, codes, dim,
                                                                         decay);
  return ctclip::check_launch("vq_ema_update");
}
