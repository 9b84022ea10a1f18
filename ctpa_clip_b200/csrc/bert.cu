// Memory-bound kernels of the injected BERT text tower (reference: ct_clip.py:685-686 calls
// transformers.BertModel(input_ids, attention_mask)[0]; BERT-base = CXR-BERT architecture, pretrained_model.py:9).
// The dense products (QKV / output / FFN projections, Q K^T, P V and their gradients) run on the tcgen05 GEMM
// (gemm_sm100.cu, batched mode for the per-head products); this file holds what sits between them:
//   embeddings gather (+ gradient scatter), masked softmax (+ attention-probability dropout) forward / backward,
//   exact-erf GELU forward / backward, hidden-state dropout + residual, bf16 column sums for bias gradients.
// One warp per row, 128-bit accesses, warp-shuffle reductions. Dropout masks come from a stateless integer hash of
// (seed, element index): the backward pass regenerates them instead of storing them.
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {
using namespace ptx;

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

// ------------------------------------------------------------------------------------------ embeddings
// x[t] = word[ids[t]] + pos[t % L] + type[0]      (BertEmbeddings.forward; token_type_ids default to 0)
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const long long* __restrict__ ids, int L, const float* __restrict__ word, const float* __restrict__ pos,
                 const float* __restrict__ type0, float* __restrict__ out, long long T, int D, int vocab) {
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  long long id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4* w = reinterpret_cast<const float4*>(word + id * D);
  const float4* p = reinterpret_cast<const float4*>(pos + (t % L) * D);
  const float4* y = reinterpret_cast<const float4*>(type0);
  float4* o = reinterpret_cast<float4*>(out + t * D);
  for (int i = lane; i < (D >> 2); i += 32) {
    const float4 a = w[i], b = p[i], c = y[i];
    o[i] = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
  }
}
// dword[ids[t]] += dx[t]; dpos[t % L] += dx[t]   (the token-type row gets a plain column sum on the host side)
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const long long* __restrict__ ids, int L, const float* __restrict__ dx, float* __restrict__ dword,
                 float* __restrict__ dpos, long long T, int D, int vocab, long long pad_id) {
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  long long id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float* g = dx + t * D;
  float* w = dword + id * D;
  float* p = dpos + (t % L) * D;
  const bool is_pad = ids[t] == pad_id;   // nn.Embedding(padding_idx): that row never receives a gradient
  for (int i = lane; i < D; i += 32) {
    const float v = g[i];
    if (!is_pad) atomicAdd(w + i, v);
    atomicAdd(p + i, v);
  }
}

// ------------------------------------------------------------------------------------------ softmax
// P = softmax(scale * S + (mask ? 0 : -inf)) over the keys of one (batch, head, query) row; Pd = dropout(P).
// S: fp32 [Z][L][L] with z = b * H + h; mask: [B][L] (1 = attend). One warp per row; lane l owns the 4 consecutive
// columns 128 j + 4 l .. + 3 of every 128-column block j (128-bit loads, 64-bit bf16 stores); kBlk = ceil(L / 128).
// L % 4 == 0 takes the vector path, anything else the scalar one (same arithmetic).
template <int kBlk>
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const float* __restrict__ S, const long long* __restrict__ mask, int H, int L, float scale,
                   __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ Pd, float p_drop, unsigned seed,
                   long long rows) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long b = row / ((long long)H * L);
  const float* s = S + row * L;
  const long long* m = mask + b * L;
  const bool vec = (L & 3) == 0;
  float v[kBlk][4];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kBlk; ++j) {
    const int c = 128 * j + 4 * lane;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec) {
      if (c < L) x = *reinterpret_cast<const float4*>(s + c);
    } else {
      if (c < L) x.x = s[c];
      if (c + 1 < L) x.y = s[c + 1];
      if (c + 2 < L) x.z = s[c + 2];
      if (c + 3 < L) x.w = s[c + 3];
    }
    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[j][e] = (c + e < L && m[c + e] != 0) ? xs[e] * scale : -INFINITY;
      mx = fmaxf(mx, v[j][e]);
    }
  }
  mx = warp_max(mx);
  if (mx == -INFINITY) mx = 0.f;  // fully masked row: all probabilities 0
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kBlk; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[j][e] = __expf(v[j][e] - mx);
      sum += v[j][e];
    }
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  const float keep_scale = 1.f / (1.f - p_drop);
#pragma unroll
  for (int j = 0; j < kBlk; ++j) {
    const int c = 128 * j + 4 * lane;
    if (c >= L) continue;
    float pr[4], pd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      pr[e] = v[j][e] * inv;
      pd[e] = (Pd != nullptr && keep_elem((unsigned long long)(row * L + c + e), seed, p_drop)) ? pr[e] * keep_scale : 0.f;
    }
    if (vec) {
      *reinterpret_cast<uint2*>(P + row * L + c) = make_uint2(pack_bf16(pr[0], pr[1]), pack_bf16(pr[2], pr[3]));
      if (Pd != nullptr)
        *reinterpret_cast<uint2*>(Pd + row * L + c) = make_uint2(pack_bf16(pd[0], pd[1]), pack_bf16(pd[2], pd[3]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c + e < L) {
          P[row * L + c + e] = __float2bfloat16_rn(pr[e]);
          if (Pd != nullptr) Pd[row * L + c + e] = __float2bfloat16_rn(pd[e]);
        }
    }
  }
}
// dS = scale * P * (dP - sum_k P_k dP_k), dP = dropout'(dPd)      (dPd fp32 [Z][L][L], dS bf16)
template <int kBlk>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dPd, int L, float scale,
                   __nv_bfloat16* __restrict__ dS, float p_drop, unsigned seed, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float keep_scale = 1.f / (1.f - p_drop);
  const bool vec = (L & 3) == 0;
  float pr[kBlk][4], g[kBlk][4];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < kBlk; ++j) {
    const int c = 128 * j + 4 * lane;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 pp = make_uint2(0u, 0u);
    if (vec) {
      if (c < L) {
        d = *reinterpret_cast<const float4*>(dPd + row * L + c);
        pp = *reinterpret_cast<const uint2*>(P + row * L + c);
      }
      pr[j][0] = bf16_lo(pp.x); pr[j][1] = bf16_hi(pp.x); pr[j][2] = bf16_lo(pp.y); pr[j][3] = bf16_hi(pp.y);
    } else {
      float* dd = &d.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pr[j][e] = 0.f;
        if (c + e < L) { dd[e] = dPd[row * L + c + e]; pr[j][e] = __bfloat162float(P[row * L + c + e]); }
      }
    }
    const float ds[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x = ds[e];
      if (p_drop > 0.f) x = keep_elem((unsigned long long)(row * L + c + e), seed, p_drop) ? x * keep_scale : 0.f;
      g[j][e] = (c + e < L) ? x : 0.f;
      dot = fmaf(pr[j][e], g[j][e], dot);
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < kBlk; ++j) {
    const int c = 128 * j + 4 * lane;
    if (c >= L) continue;
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = scale * pr[j][e] * (g[j][e] - dot);
    if (vec) {
      *reinterpret_cast<uint2*>(dS + row * L + c) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c + e < L) dS[row * L + c + e] = __float2bfloat16_rn(o[e]);
    }
  }
}

// ------------------------------------------------------------------------------------------ GELU (erf)
__device__ __forceinline__ float gelu_f(float x) { return gelu_fast(x); }
__device__ __forceinline__ float gelu_df(float x) { return gelu_grad_fast(x); }

__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const uint4* __restrict__ h, uint4* __restrict__ out, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = h[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = pack_bf16(gelu_f(bf16_lo(w[k])), gelu_f(bf16_hi(w[k])));
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const uint4* __restrict__ h, const uint4* __restrict__ dy, uint4* __restrict__ dh, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = h[i], g = dy[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w}, d[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = pack_bf16(gelu_df(bf16_lo(w[k])) * bf16_lo(d[k]), gelu_df(bf16_hi(w[k])) * bf16_hi(d[k]));
    dh[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------ dropout (+ residual)
// out = dropout(y) (+ resid); also the backward of dropout (resid = null, y = upstream gradient)
__global__ void __launch_bounds__(256)
dropout_add_kernel(const float4* __restrict__ y, const float4* __restrict__ resid, float4* __restrict__ out, long long nvec,
                   float p_drop, unsigned seed) {
  const float ks = 1.f / (1.f - p_drop);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float4 v = y[i];
    v.x = keep_elem(4ull * i + 0, seed, p_drop) ? v.x * ks : 0.f;
    v.y = keep_elem(4ull * i + 1, seed, p_drop) ? v.y * ks : 0.f;
    v.z = keep_elem(4ull * i + 2, seed, p_drop) ? v.z * ks : 0.f;
    v.w = keep_elem(4ull * i + 3, seed, p_drop) ? v.w * ks : 0.f;
    if (resid != nullptr) {
      const float4 r = resid[i];
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------ bf16 column sums
// out[c] += sum_rows x[row][c]   (bias gradients of the bf16 activation gradients); ld = row pitch in elements
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int dim, long long ld, float* __restrict__ out,
                   int rows_per_block) {
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x * 2; c < dim; c += blockDim.x * 2) {
    float a = 0.f, b = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const uint32_t u = *reinterpret_cast<const uint32_t*>(x + r * ld + c);
      a += bf16_lo(u);
      b += bf16_hi(u);
    }
    atomicAdd(out + c, a);
    if (c + 1 < dim) atomicAdd(out + c + 1, b);
  }
}

__global__ void __launch_bounds__(256)
colsum_bf16_tile_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int dim, long long ld, float* __restrict__ out,
                        int rows_per_block) {
  colsum_tile<__nv_bfloat16>(x, rows, dim, ld, out, rows_per_block);
}

int grid_for(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  const long long cap = (long long)ctclip::sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int ctclip_bert_embed_fwd(const long long* ids, long long tokens, int seq_len, const float* word,
                                     const float* pos, const float* type0, int dim, int vocab, float* out, void* stream) {
  if (tokens <= 0) return CTCLIP_OK;
  if (dim % 4 || seq_len <= 0) return ctclip::fail(CTCLIP_E_SHAPE, "bert_embed_fwd: dim must be a multiple of 4");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  embed_fwd_kernel<<<(unsigned)((tokens + 7) / 8), 256, 0, (cudaStream_t)stream>>>(ids, seq_len, word, pos, type0, out,
                                                                                   tokens, dim, vocab);
  return ctclip::check_launch("bert_embed_fwd");
}

extern "C" int ctclip_bert_embed_bwd(const long long* ids, long long tokens, int seq_len, const float* dx, int dim,
                                     int vocab, long long pad_id, float* dword, float* dpos, void* stream) {
  if (tokens <= 0) return CTCLIP_OK;
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  embed_bwd_kernel<<<(unsigned)((tokens + 7) / 8), 256, 0, (cudaStream_t)stream>>>(ids, seq_len, dx, dword, dpos, tokens,
                                                                                   dim, vocab, pad_id);
  return ctclip::check_launch("bert_embed_bwd");
}

#define SOFTMAX_DISPATCH(KERNEL, ...)                                                         \
  do {                                                                                        \
    const int blk = (seq_len + 127) / 128;                                                    \
    const unsigned blocks = (unsigned)((rows + 7) / 8);                                       \
    if (blk <= 1) KERNEL<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);           \
    else if (blk <= 2) KERNEL<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);      \
    else if (blk <= 4) KERNEL<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);      \
    else KERNEL<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);                    \
  } while (0)

extern "C" int ctclip_bert_softmax_fwd(const float* scores, const long long* mask, int batch, int heads, int seq_len,
                                       float scale, void* probs, void* probs_dropped, float p_drop, unsigned seed,
                                       void* stream) {
  if (batch <= 0 || heads <= 0 || seq_len <= 0) return ctclip::fail(CTCLIP_E_SHAPE, "bert_softmax_fwd: empty problem");
  if (seq_len > 1024) return ctclip::fail(CTCLIP_E_SHAPE, "bert_softmax_fwd: sequences longer than 1024 are not supported");
  if (p_drop < 0.f || p_drop >= 1.f) return ctclip::fail(CTCLIP_E_SHAPE, "bert_softmax_fwd: dropout must be in [0, 1)");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  const long long rows = (long long)batch * heads * seq_len;
  SOFTMAX_DISPATCH(softmax_fwd_kernel, scores, mask, heads, seq_len, scale, (__nv_bfloat16*)probs,
                   (__nv_bfloat16*)(p_drop > 0.f ? probs_dropped : nullptr), p_drop, seed, rows);
  return ctclip::check_launch("bert_softmax_fwd");
}

extern "C" int ctclip_bert_softmax_bwd(const void* probs, const float* dprobs_dropped, int batch, int heads, int seq_len,
                                       float scale, void* dscores, float p_drop, unsigned seed, void* stream) {
  if (batch <= 0 || heads <= 0 || seq_len <= 0) return ctclip::fail(CTCLIP_E_SHAPE, "bert_softmax_bwd: empty problem");
  if (seq_len > 1024) return ctclip::fail(CTCLIP_E_SHAPE, "bert_softmax_bwd: sequences longer than 1024 are not supported");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  const long long rows = (long long)batch * heads * seq_len;
  SOFTMAX_DISPATCH(softmax_bwd_kernel, (const __nv_bfloat16*)probs, dprobs_dropped, seq_len, scale,
                   (__nv_bfloat16*)dscores, p_drop, seed, rows);
  return ctclip::check_launch("bert_softmax_bwd");
}

extern "C" int ctclip_gelu_fwd(const void* h, void* out, long long n, void* stream) {
  if (n <= 0) return CTCLIP_OK;
  if (n % 8) return ctclip::fail(CTCLIP_E_ALIGN, "gelu_fwd: element count must be a multiple of 8");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  gelu_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)h, (uint4*)out, n / 8);
  return ctclip::check_launch("gelu_fwd");
}

extern "C" int ctclip_gelu_bwd(const void* h, const void* dy, void* dh, long long n, void* stream) {
  if (n <= 0) return CTCLIP_OK;
  if (n % 8) return ctclip::fail(CTCLIP_E_ALIGN, "gelu_bwd: element count must be a multiple of 8");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  gelu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)h, (const uint4*)dy, (uint4*)dh,
                                                                         n / 8);
  return ctclip::check_launch("gelu_bwd");
}

extern "C" int ctclip_dropout_add(const float* y, const float* resid, float* out, long long n, float p_drop,
                                  unsigned seed, void* stream) {
  if (n <= 0) return CTCLIP_OK;
  if (n % 4) return ctclip::fail(CTCLIP_E_ALIGN, "dropout_add: element count must be a multiple of 4");
  if (p_drop < 0.f || p_drop >= 1.f) return ctclip::fail(CTCLIP_E_SHAPE, "dropout_add: dropout must be in [0, 1)");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  dropout_add_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)y, (const float4*)resid,
                                                                            (float4*)out, n / 4, p_drop, seed);
  return ctclip::check_launch("dropout_add");
}

extern "C" int ctclip_colsum_bf16(const void* x, long long rows, int dim, long long ld, float* out, void* stream) {
  if (rows <= 0 || dim <= 0) return CTCLIP_OK;
  if ((dim % 2) || (ld % 2)) return ctclip::fail(CTCLIP_E_ALIGN, "colsum_bf16: dim and ld must be even");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  if (dim % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int bx = (dim / 8 + 255) / 256;
    long long rpb8 = rows * bx / (2LL * ctclip::sm_count());   // about two CTAs per SM
    rpb8 = rpb8 < 8 ? 8 : (rpb8 > 64 ? 64 : rpb8 / 8 * 8);
    dim3 grid((unsigned)bx, (unsigned)((rows + rpb8 - 1) / rpb8));
    colsum_bf16_tile_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, dim, ld, out, (int)rpb8);
    return ctclip::check_launch("colsum_bf16");
  }
  int rpb = 8;   // rows per block: enough blocks to fill the chip even for a few thousand rows
  long long blocks = (rows + rpb - 1) / rpb;
  const long long cap = (long long)ctclip::sm_count() * 16;
  if (blocks > cap) {
    rpb = (int)((rows + cap - 1) / cap);
    blocks = (rows + rpb - 1) / rpb;
  }
  colsum_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, dim, ld, out, rpb);
  return ctclip::check_launch("colsum_bf16");
}
