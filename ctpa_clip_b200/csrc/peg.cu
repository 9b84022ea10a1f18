// PEG: depth-wise causal 3x3x3 convolution over the token grid + residual (reference attention.py:56-84).
//
// Tokens stay in ONE canonical HBM layout (b, t, h, w, d) for the whole encoder; channels-last, so every tap is a
// coalesced 128-bit load across channels. The reference feeds the temporal transformer a '(b h w) t d' tensor and
// PEG *reshapes* (does not permute) it to (b, t, h, w, d) (attention.py:69-70 with ctvit.py:325-327): the stencil
// then runs over a re-interpretation of the flat (h, w, t) order. `temporal=1` reproduces exactly that
// re-interpretation through index arithmetic (valid for any t, h, w, not only the cubic production grid).
#include "ptx.cuh"
#include "ctclip_internal.h"
#include <type_traits>
#include <cstdlib>

namespace {

struct PegGrid {
  int t, h, w, temporal;
};

__device__ __forceinline__ void canon_to_virtual(int c, const PegGrid& g, int& vt, int& vh, int& vw) {
  int f = c;
  if (g.temporal) {
    const int wi = c % g.w, hi = (c / g.w) % g.h, ti = c / (g.w * g.h);
    f = (hi * g.w + wi) * g.t + ti;
  }
  vw = f % g.w;
  vh = (f / g.w) % g.h;
  vt = f / (g.w * g.h);
}
__device__ __forceinline__ int virtual_to_canon(int vt, int vh, int vw, const PegGrid& g) {
  const int f = (vt * g.h + vh) * g.w + vw;
  if (!g.temporal) return f;
  const int ti = f % g.t, wi = (f / g.t) % g.w, hi = f / (g.t * g.w);
  return (ti * g.h + hi) * g.w + wi;
}

// out[c] = in[c] + (bias) + sum_taps w27[tap] * in[nbr(c, tap)]
//   flip = 0: forward stencil, neighbour offsets (kt-2, kh-1, kw-1)
//   flip = 1: transposed stencil (gradient w.r.t. the input), offsets (2-kt, 1-kh, 1-kw), no bias
__global__ void __launch_bounds__(256)
peg_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w27,
           const float* __restrict__ bias, int batch, PegGrid g, int dim, int flip,
           __nv_bfloat16* __restrict__ out_bf16) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int per_b = g.t * g.h * g.w;
  const long long token = (long long)blockIdx.x * 8 + warp;
  if (token >= (long long)batch * per_b) return;
  const int v4 = blockIdx.y * 32 + lane;  // float4 channel index
  if (v4 >= (dim >> 2)) return;
  const int b = (int)(token / per_b);
  const int c = (int)(token - (long long)b * per_b);
  int vt, vh, vw;
  canon_to_virtual(c, g, vt, vh, vw);
  const float4* base = reinterpret_cast<const float4*>(in + (long long)b * per_b * dim);
  const int dv = dim >> 2;
  float4 acc = base[(long long)c * dv + v4];
  if (!flip && bias != nullptr) {
    const float4 bb = reinterpret_cast<const float4*>(bias)[v4];
    acc.x += bb.x; acc.y += bb.y; acc.z += bb.z; acc.w += bb.w;
  }
#pragma unroll
  for (int kt = 0; kt < 3; ++kt) {
    const int nt = vt + (flip ? 2 - kt : kt - 2);
    if (nt < 0 || nt >= g.t) continue;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int nh = vh + (flip ? 1 - kh : kh - 1);
      if (nh < 0 || nh >= g.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int nw = vw + (flip ? 1 - kw : kw - 1);
        if (nw < 0 || nw >= g.w) continue;
        const int nc = virtual_to_canon(nt, nh, nw, g);
        const float4 xv = base[(long long)nc * dv + v4];
        const float4 wv = reinterpret_cast<const float4*>(w27 + ((kt * 3 + kh) * 3 + kw) * dim)[v4];
        acc.x = fmaf(wv.x, xv.x, acc.x); acc.y = fmaf(wv.y, xv.y, acc.y);
        acc.z = fmaf(wv.z, xv.z, acc.z); acc.w = fmaf(wv.w, xv.w, acc.w);
      }
    }
  }
  reinterpret_cast<float4*>(out)[token * dv + v4] = acc;
  if (out_bf16 != nullptr)
    reinterpret_cast<uint2*>(out_bf16)[token * dv + v4] =
        make_uint2(ptx::pack_bf16(acc.x, acc.y), ptx::pack_bf16(acc.z, acc.w));
}

// dw27[tap][ch] += sum_tokens dy[c][ch] * x[nbr(c,tap)][ch] ; dbias[ch] += sum_tokens dy[c][ch]
__global__ void __launch_bounds__(128)
peg_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw27,
                 float* __restrict__ dbias, int batch, PegGrid g, int dim, int tokens_per_block) {
  const int v4 = blockIdx.y * blockDim.x + threadIdx.x;
  const int dv = dim >> 2;
  if (v4 >= dv) return;
  const int per_b = g.t * g.h * g.w;
  const long long total = (long long)batch * per_b;
  const long long t0 = (long long)blockIdx.x * tokens_per_block;
  const long long t1 = min(total, t0 + tokens_per_block);
  float4 acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 accb = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long token = t0; token < t1; ++token) {
    const int b = (int)(token / per_b);
    const int c = (int)(token - (long long)b * per_b);
    int vt, vh, vw;
    canon_to_virtual(c, g, vt, vh, vw);
    const float4 d = reinterpret_cast<const float4*>(dy)[token * dv + v4];
    accb.x += d.x; accb.y += d.y; accb.z += d.z; accb.w += d.w;
    const float4* base = reinterpret_cast<const float4*>(x + (long long)b * per_b * dim);
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const int nt = vt + kt - 2;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int nh = vh + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int nw = vw + kw - 1;
          if (nt < 0 || nt >= g.t || nh < 0 || nh >= g.h || nw < 0 || nw >= g.w) continue;
          const int nc = virtual_to_canon(nt, nh, nw, g);
          const float4 xv = base[(long long)nc * dv + v4];
          float4& a = acc[(kt * 3 + kh) * 3 + kw];
          a.x = fmaf(d.x, xv.x, a.x); a.y = fmaf(d.y, xv.y, a.y);
          a.z = fmaf(d.z, xv.z, a.z); a.w = fmaf(d.w, xv.w, a.w);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 27; ++i) {
    float* p = dw27 + i * dim + 4 * v4;
    atomicAdd(p + 0, acc[i].x); atomicAdd(p + 1, acc[i].y); atomicAdd(p + 2, acc[i].z); atomicAdd(p + 3, acc[i].w);
  }
  if (dbias != nullptr) {
    float* p = dbias + 4 * v4;
    atomicAdd(p + 0, accb.x); atomicAdd(p + 1, accb.y); atomicAdd(p + 2, accb.z); atomicAdd(p + 3, accb.w);
  }
}


// ------------------------------------------------------------------------------------------------ tiled kernels
// A CTA owns a (TH x TW) patch of the virtual (vh, vw) plane for one 32-channel slice and walks the whole vt axis with a
// rolling window of three input planes in shared memory: every input element is read ~1.5x from HBM/L2 instead of 27x.
// lane = channel (its 27 taps stay in registers), warp = output position. MODE 0: y = x + conv(x) + bias;
// MODE 1: dx = dy + conv^T(dy) (walks vt downwards, taps mirrored); MODE 2: dw27 / dbias accumulation.
constexpr int TH = 8, TW = 8, PH = TH + 2, PW = TW + 2;

template <int MODE>
__global__ void __launch_bounds__(256)
peg_tiled_kernel(const float* __restrict__ in, const float* __restrict__ in2, float* __restrict__ out,
                 __nv_bfloat16* __restrict__ out_bf16, const float* __restrict__ w27, const float* __restrict__ bias,
                 float* __restrict__ dw27, float* __restrict__ dbias, PegGrid g, int dim) {
  __shared__ float planes[3][PH * PW][32];
  __shared__ float dyp[MODE == 2 ? TH * TW : 1][32];
  __shared__ int tok[PH * PW];  // canonical token index of every position of the current plane (-1 outside the grid)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = blockIdx.y * 32 + lane;
  const int tiles_w = (g.w + TW - 1) / TW;
  const int h0 = (blockIdx.x / tiles_w) * TH, w0 = (blockIdx.x % tiles_w) * TW;
  const int b = blockIdx.z;
  const long long per_b = (long long)g.t * g.h * g.w;
  const float* base = in + b * per_b * dim;
  const bool flip = (MODE == 1);

  float wt[27];
  if (MODE != 2) {
#pragma unroll
    for (int kt = 0; kt < 3; ++kt)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          // age a = 2 - kt planes behind the cursor; spatial taps mirrored for the transposed stencil
          const int src = flip ? ((kt * 3 + (2 - kh)) * 3 + (2 - kw)) : ((kt * 3 + kh) * 3 + kw);
          wt[(kt * 3 + kh) * 3 + kw] = w27[src * dim + ch];
        }
  } else {
#pragma unroll
    for (int i = 0; i < 27; ++i) wt[i] = 0.f;
  }
  float bsum = 0.f;
  const float bval = (MODE == 0 && bias != nullptr) ? bias[ch] : 0.f;

  for (int step = 0; step < g.t; ++step) {
    const int vt = flip ? g.t - 1 - step : step;
    // ---- token indices of this plane (one thread per position: the divisions happen once, not per warp)
    if (threadIdx.x < PH * PW) {
      const int vh = h0 - 1 + threadIdx.x / PW, vw = w0 - 1 + threadIdx.x % PW;
      tok[threadIdx.x] = (vh >= 0 && vh < g.h && vw >= 0 && vw < g.w) ? virtual_to_canon(vt, vh, vw, g) : -1;
    }
    __syncthreads();
    // ---- load plane vt into slot step % 3 (zero outside the grid)
    float (*slot)[32] = planes[step % 3];
    for (int pos = warp; pos < PH * PW; pos += 8) {
      const int tk = tok[pos];
      slot[pos][lane] = (tk >= 0) ? base[(long long)tk * dim + ch] : 0.f;
    }
    if (MODE == 2) {
      const float* base2 = in2 + b * per_b * dim;
      for (int pos = warp; pos < TH * TW; pos += 8) {
        const int tk = tok[(pos / TW + 1) * PW + pos % TW + 1];
        dyp[pos][lane] = (tk >= 0) ? base2[(long long)tk * dim + ch] : 0.f;
      }
    }
    __syncthreads();
    // ---- outputs of plane vt: warp = output row, lane = channel; each staged value feeds up to 3 outputs of the row
    {
      const int oh = warp;  // TH == 8 == warps per CTA
      float acc[TW];
      float dcur[TW];
#pragma unroll
      for (int ow = 0; ow < TW; ++ow) {
        acc[ow] = planes[step % 3][(oh + 1) * PW + ow + 1][lane] + bval;  // residual + bias (MODE 0/1)
        dcur[ow] = (MODE == 2) ? dyp[oh * TW + ow][lane] : 0.f;
        if (MODE == 2) bsum += dcur[ow];
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (step - a < 0) continue;
        float (*pl)[32] = planes[(step - a) % 3];
        const int kt = 2 - a;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          float in[PW];
#pragma unroll
          for (int x = 0; x < PW; ++x) in[x] = pl[(oh + kh) * PW + x][lane];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int wi = (kt * 3 + kh) * 3 + kw;
            if (MODE != 2) {
#pragma unroll
              for (int ow = 0; ow < TW; ++ow) acc[ow] = fmaf(wt[wi], in[ow + kw], acc[ow]);
            } else {
#pragma unroll
              for (int ow = 0; ow < TW; ++ow) wt[wi] = fmaf(dcur[ow], in[ow + kw], wt[wi]);
            }
          }
        }
      }
      if (MODE != 2) {
#pragma unroll
        for (int ow = 0; ow < TW; ++ow) {
          const int tk = tok[(oh + 1) * PW + ow + 1];
          if (tk < 0) continue;
          const long long o = (b * per_b + tk) * dim + ch;
          out[o] = acc[ow];
          if (MODE == 1 && out_bf16 != nullptr) out_bf16[o] = __float2bfloat16_rn(acc[ow]);
        }
      }
    }
    __syncthreads();
  }
  if (MODE == 2) {
    // reduce the 8 warps through shared memory (reuse the plane buffers), one atomic per (tap, channel) per CTA
    float (*red)[32] = planes[0];  // [8*28][32] fits in 3*100 rows
#pragma unroll
    for (int i = 0; i < 27; ++i) red[warp * 28 + i][lane] = wt[i];
    red[warp * 28 + 27][lane] = bsum;
    __syncthreads();
    for (int i = warp; i < 28; i += 8) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w * 28 + i][lane];
      if (i < 27) atomicAdd(dw27 + i * dim + ch, s);
      else if (dbias != nullptr) atomicAdd(dbias + ch, s);
    }
  }
}


// ------------------------------------------------------------------------------------------------ marching kernels
// Fast path for grids whose virtual axes are affine in the canonical token index (every spatial call; the temporal call
// when t == h == w, i.e. the production 24^3 grid): token(vt, vh, vw) = vt*st + vh*sh + vw*sw, with w % MTW == 0.
// A warp owns one output row segment of MTW positions x 64 channels (lane = channel pair, 256-byte coalesced token
// slices) and marches along vt with THREE rolling accumulator rows in registers: input plane p is loaded once
// (3 rows x (MTW+2) positions, straight from L1/L2 — no shared memory, no barriers) and scattered into the outputs
// p, p+1, p+2 it feeds: 4.5 LDG.64 + 27 FMA x 2 channels per output position. Out-of-grid rows are handled by zeroing the
// nine taps of that row (row validity is per-warp constant) and clamping the address; the two halo columns by a select.
// Neighbouring warps / CTAs share halo rows through L1 / L2, so HBM sees every input exactly once.
// MODE 0: y = x + conv(x) + bias; MODE 1: dx = dy + conv^T(dy) (marches downwards, spatial taps mirrored);
// MODE 2: dw27 / dbias accumulation (x planes march, three dy rows roll).
constexpr int MTW = 4, MTH = 8, MCH = 64;
// STAGE (default): the (MTH + 2) x (MTW + 2) x 64-channel neighbourhood of an input plane travels global -> shared memory with
// cp.async, two planes ahead of the one being consumed (three-slot ring, ONE barrier per plane), and the warps read their
// three rows from shared memory. ncu on the direct-load version: 5-6.7 long-scoreboard stall cycles per issued instruction
// at 16 warps / SM (128 registers) — the 18 LDG.64 of a plane are consumed right after they are issued — 35-41 % DRAM.
constexpr int PEG_STAGES = 3;
constexpr int PEG_TILE_POS = (MTH + 2) * (MTW + 2);            // positions of one staged plane tile
constexpr int PEG_TILE_BYTES = PEG_TILE_POS * MCH * 4;         // 15360
constexpr int PEG_CHUNKS = PEG_TILE_BYTES / 16;                // 960 16-byte chunks: <= 4 per thread

__device__ __forceinline__ void peg_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void peg_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void peg_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float2 peg_lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

template <int MODE, bool STAGE>
__global__ void __launch_bounds__(256, 2)
peg_march_kernel(const float* __restrict__ in, const float* __restrict__ in2, float* __restrict__ out,
                 __nv_bfloat16* __restrict__ out_bf16, const float* __restrict__ w27, const float* __restrict__ bias,
                 float* __restrict__ dw27, float* __restrict__ dbias, int T, int H, int W, int st, int sh, int sw,
                 int dim) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = blockIdx.y * MCH + 2 * lane;
  const int tiles_w = W / MTW;
  const int h0 = (blockIdx.x / tiles_w) * MTH, w0 = (blockIdx.x % tiles_w) * MTW;
  const int oh = h0 + warp;
  const long long vol = (long long)blockIdx.z * T * H * W * dim;
  const bool row_ok = oh < H;
  constexpr bool flip = (MODE == 1);
  const bool left_ok = w0 > 0, right_ok = w0 + MTW < W;
  bool rv[3];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) rv[kh] = row_ok && oh + kh - 1 >= 0 && oh + kh - 1 < H;

  float2 wt[27];
#pragma unroll
  for (int kt = 0; kt < 3; ++kt)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int src = flip ? ((kt * 3 + (2 - kh)) * 3 + (2 - kw)) : ((kt * 3 + kh) * 3 + kw);
        float2 w = make_float2(0.f, 0.f);
        if (MODE != 2 && rv[kh]) w = __ldg(reinterpret_cast<const float2*>(w27 + src * dim + ch));
        wt[(kt * 3 + kh) * 3 + kw] = w;
      }
  float2 bval = make_float2(0.f, 0.f);
  if (MODE == 0 && bias != nullptr) bval = __ldg(reinterpret_cast<const float2*>(bias + ch));
  float2 bsum = make_float2(0.f, 0.f);

  // BYTE offsets of the 3 x (MTW+2) neighbourhood inside plane 0 (rows / halo columns clamped into the grid):
  // roff is per thread (row, channel pair), coff is CTA-uniform (column) -> address = uniform base + roff + coff
  unsigned roff[3], coff[MTW + 2];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) roff[kh] = (unsigned)((min(max(oh + kh - 1, 0), H - 1) * sh * dim + ch) * 4);
#pragma unroll
  for (int x = 0; x < MTW + 2; ++x) coff[x] = (unsigned)(min(max(w0 + x - 1, 0), W - 1) * sw * dim * 4);
  const long long pstride = (long long)st * dim * 4;   // bytes
  const char* base = reinterpret_cast<const char*>(in + vol);
  const char* base2 = reinterpret_cast<const char*>(in2 + vol);
  char* obase = reinterpret_cast<char*>(out + vol);
  char* obase_bf = reinterpret_cast<char*>(out_bf16 + vol);

  float2 acc[3][MTW];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int ow = 0; ow < MTW; ++ow) acc[s][ow] = make_float2(0.f, 0.f);

  // ---- staging: this thread's <= 4 chunks of a plane tile (same offsets for every plane)
  extern __shared__ __align__(16) uint8_t peg_smem[];
  const uint32_t ring = ptx::smem_u32(peg_smem);
  unsigned g_off[4];     // byte offset of the chunk inside plane 0 (row / column clamped into the grid)
  uint32_t s_off[4];     // byte offset inside a ring slot
  bool c_use[4];
  if constexpr (STAGE) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = threadIdx.x + u * 256;
      c_use[u] = c < PEG_CHUNKS;
      const int pos = min(c, PEG_CHUNKS - 1) >> 4, q = c & 15;      // 16 chunks = the 64 channels of one position
      const int r = pos / (MTW + 2), x = pos - r * (MTW + 2);
      g_off[u] = (unsigned)((min(max(h0 + r - 1, 0), H - 1) * sh + min(max(w0 + x - 1, 0), W - 1) * sw) * dim +
                            blockIdx.y * MCH + 4 * q) * 4u;
      s_off[u] = (uint32_t)(pos * MCH * 4 + q * 16);
    }
  }
  auto stage_plane = [&](int p, int slot) {        // plane p (already a valid plane index) -> ring slot
    const char* pb = base + p * pstride;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c_use[u]) peg_cp_async16(ring + slot * PEG_TILE_BYTES + s_off[u], pb + g_off[u]);
  };
  if constexpr (STAGE) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {                  // planes of steps 0 and 1
      if (s < T) stage_plane(flip ? T - 1 - s : s, s);
      peg_cp_commit();
    }
  }

  if (MODE == 2 && row_ok) {  // dy rows of planes 0 and 1
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int ow = 0; ow < MTW; ++ow)
        if (s < T) acc[s][ow] = __ldg(reinterpret_cast<const float2*>(base2 + s * pstride + roff[1] + coff[ow + 1]));
  }

  auto body = [&](auto phc, int step) {
    constexpr int PH = decltype(phc)::value;
    const int p = flip ? T - 1 - step : step;
    const char* pb = base + p * pstride;
    uint32_t srow = 0;
    if constexpr (STAGE) {
      peg_cp_wait<1>();                             // this thread's chunks of plane `step` have landed (step + 1 may be pending)
      __syncthreads();                              // ... everybody's; and every warp is done reading slot (step - 1) % 3
      if (step + 2 < T) stage_plane(flip ? T - 3 - step : step + 2, (step + 2) % PEG_STAGES);
      peg_cp_commit();
      srow = ring + (step % PEG_STAGES) * PEG_TILE_BYTES + (uint32_t)(warp * (MTW + 2)) * (MCH * 4) + lane * 8;
    }
    if (MODE == 2) {
      // dy row of plane p + 2 replaces the slot that held plane p - 1
      const char* b2 = base2 + (p + 2) * pstride;
#pragma unroll
      for (int ow = 0; ow < MTW; ++ow)
        acc[(PH + 2) % 3][ow] = (row_ok && p + 2 < T) ? __ldg(reinterpret_cast<const float2*>(b2 + roff[1] + coff[ow + 1]))
                                                      : make_float2(0.f, 0.f);
#pragma unroll
      for (int ow = 0; ow < MTW; ++ow) { bsum.x += acc[PH][ow].x; bsum.y += acc[PH][ow].y; }
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      float2 v[MTW + 2];
      const char* rp = pb + roff[kh];
#pragma unroll
      for (int x = 0; x < MTW + 2; ++x) {
        if constexpr (STAGE) v[x] = peg_lds64(srow + (uint32_t)((kh * (MTW + 2) + x) * (MCH * 4)));
        else v[x] = __ldg(reinterpret_cast<const float2*>(rp + coff[x]));
      }
      if (!left_ok) v[0] = make_float2(0.f, 0.f);
      if (!right_ok) v[MTW + 1] = make_float2(0.f, 0.f);
      if (MODE != 2 && kh == 1) {  // residual: the centre tap of the plane that completes at this step
#pragma unroll
        for (int ow = 0; ow < MTW; ++ow) {
          acc[PH][ow].x += v[ow + 1].x + bval.x;
          acc[PH][ow].y += v[ow + 1].y + bval.y;
        }
      }
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int slot = (PH + 2 - kt) % 3;  // plane p feeds output p + 2 - kt (MODE 2: pairs with dy of that plane)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int wi = (kt * 3 + kh) * 3 + kw;
#pragma unroll
          for (int ow = 0; ow < MTW; ++ow) {
            if (MODE != 2) acc[slot][ow] = ptx::ffma2(wt[wi], v[ow + kw], acc[slot][ow]);
            else wt[wi] = ptx::ffma2(acc[slot][ow], v[ow + kw], wt[wi]);
          }
        }
      }
    }
    if (MODE != 2) {
      if (row_ok) {
        char* ob = obase + p * pstride + roff[1];
        char* obb = obase_bf + p * (pstride >> 1) + (roff[1] >> 1);
#pragma unroll
        for (int ow = 0; ow < MTW; ++ow) {
          *reinterpret_cast<float2*>(ob + coff[ow + 1]) = acc[PH][ow];
          if (MODE == 1 && out_bf16 != nullptr)
            *reinterpret_cast<uint32_t*>(obb + (coff[ow + 1] >> 1)) = ptx::pack_bf16(acc[PH][ow].x, acc[PH][ow].y);
        }
      }
#pragma unroll
      for (int ow = 0; ow < MTW; ++ow) acc[PH][ow] = make_float2(0.f, 0.f);
    }
  };

  int step = 0;
  for (; step + 3 <= T; step += 3) {
    body(std::integral_constant<int, 0>{}, step);
    body(std::integral_constant<int, 1>{}, step + 1);
    body(std::integral_constant<int, 2>{}, step + 2);
  }
  if (step < T) body(std::integral_constant<int, 0>{}, step);
  if (step + 1 < T) body(std::integral_constant<int, 1>{}, step + 1);

  if (MODE == 2) {
    __shared__ float red[8 * 28][32];   // the two channels of a lane are reduced one after the other (28 KB)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int kt = 0; kt < 3; ++kt)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int i = (kt * 3 + kh) * 3 + kw;
            red[warp * 28 + i][lane] = rv[kh] ? (c ? wt[i].y : wt[i].x) : 0.f;   // clamped rows contributed garbage
          }
      red[warp * 28 + 27][lane] = c ? bsum.y : bsum.x;
      __syncthreads();
      for (int i = warp; i < 28; i += 8) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * 28 + i][lane];
        if (i < 27) atomicAdd(dw27 + i * dim + ch + c, s);
        else if (dbias != nullptr) atomicAdd(dbias + ch + c, s);
      }
      __syncthreads();
    }
  }
}

// virtual-axis token strides when the mapping is affine; false otherwise
bool affine_strides(int t, int h, int w, int temporal, int& st, int& sh, int& sw) {
  if (!temporal) { st = h * w; sh = w; sw = 1; return true; }
  if (t == h && h == w) { st = w; sh = 1; sw = h * w; return true; }   // virtual (vt, vh, vw) = canonical (hi, wi, ti)
  return false;
}

template <int MODE>
int launch_tiled(const float* in, const float* in2, float* out, void* out_bf16, const float* w27, const float* bias,
                 float* dw27, float* dbias, int batch, int t, int h, int w, int dim, int temporal, void* stream,
                 const char* what) {
  int st, sh, sw;
  if (affine_strides(t, h, w, temporal, st, sh, sw) && (long long)t * h * w * dim < (1ll << 31) && w % MTW == 0 &&
      dim % MCH == 0) {
    dim3 grid((unsigned)(((h + MTH - 1) / MTH) * (w / MTW)), (unsigned)(dim / MCH), (unsigned)batch);
    const char* e = getenv("CTCLIP_PEG_STAGE");    // 0: direct global loads (A/B arm)
    if (!(e != nullptr && e[0] == '0')) {
      constexpr int smem = PEG_STAGES * PEG_TILE_BYTES;
      static bool configured = false;              // per MODE instantiation
      if (!configured) {
        if (cudaFuncSetAttribute(peg_march_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
          return ctclip::fail(CTCLIP_E_CUDA, "%s: cudaFuncSetAttribute", what);
        configured = true;
      }
      peg_march_kernel<MODE, true><<<grid, 256, smem, (cudaStream_t)stream>>>(in, in2, out, (__nv_bfloat16*)out_bf16, w27, bias,
                                                                              dw27, dbias, t, h, w, st, sh, sw, dim);
    } else {
      peg_march_kernel<MODE, false><<<grid, 256, 0, (cudaStream_t)stream>>>(in, in2, out, (__nv_bfloat16*)out_bf16, w27, bias,
                                                                            dw27, dbias, t, h, w, st, sh, sw, dim);
    }
    return ctclip::check_launch(what);
  }
  PegGrid g{t, h, w, temporal};
  int vt = t, vh = h, vw = w;  // the virtual grid has the same extents (reshape, not permute)
  (void)vt;
  dim3 grid((unsigned)(((vh + TH - 1) / TH) * ((vw + TW - 1) / TW)), (unsigned)(dim / 32), (unsigned)batch);
  peg_tiled_kernel<MODE><<<grid, 256, 0, (cudaStream_t)stream>>>(in, in2, out, (__nv_bfloat16*)out_bf16, w27, bias, dw27,
                                                                dbias, g, dim);
  return ctclip::check_launch(what);
}

int check(int batch, int t, int h, int w, int dim, const char* what) {
  if (batch <= 0 || t <= 0 || h <= 0 || w <= 0) return ctclip::fail(CTCLIP_E_SHAPE, "%s: empty grid", what);
  if (dim % 4 || dim <= 0) return ctclip::fail(CTCLIP_E_SHAPE, "%s: dim must be a multiple of 4", what);
  return ctclip::require_sm100();
}

}  // namespace

extern "C" int ctclip_peg_fwd(const float* x, float* y, const float* w27, const float* bias, int batch, int t, int h,
                              int w, int dim, int temporal, void* stream) {
  int rc = check(batch, t, h, w, dim, "peg_fwd");
  if (rc) return rc;
  if (dim % 32 == 0 && batch <= 65535)
    return launch_tiled<0>(x, nullptr, y, nullptr, w27, bias, nullptr, nullptr, batch, t, h, w, dim, temporal, stream, "peg_fwd");
  const long long tokens = (long long)batch * t * h * w;
  dim3 grid((unsigned)((tokens + 7) / 8), (unsigned)((dim / 4 + 31) / 32));
  peg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, w27, bias, batch, PegGrid{t, h, w, temporal}, dim, 0,
                                                     nullptr);
  return ctclip::check_launch("peg_fwd");
}

// dx = dy + conv^T(dy); optional bf16 copy of dx
extern "C" int ctclip_peg_bwd_data(const float* dy, float* dx, void* dx_bf16, const float* w27, int batch, int t, int h,
                                   int w, int dim, int temporal, void* stream) {
  int rc = check(batch, t, h, w, dim, "peg_bwd_data");
  if (rc) return rc;
  if (dim % 32 == 0 && batch <= 65535)
    return launch_tiled<1>(dy, nullptr, dx, dx_bf16, w27, nullptr, nullptr, nullptr, batch, t, h, w, dim, temporal, stream,
                           "peg_bwd_data");
  const long long tokens = (long long)batch * t * h * w;
  dim3 grid((unsigned)((tokens + 7) / 8), (unsigned)((dim / 4 + 31) / 32));
  peg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, dx, w27, nullptr, batch, PegGrid{t, h, w, temporal}, dim, 1,
                                                     (__nv_bfloat16*)dx_bf16);
  return ctclip::check_launch("peg_bwd_data");
}

// dw27 += ..., dbias += ...   (x = the PEG input, dy = gradient of the PEG output)
extern "C" int ctclip_peg_bwd_weight(const float* x, const float* dy, float* dw27, float* dbias, int batch, int t, int h,
                                     int w, int dim, int temporal, void* stream) {
  int rc = check(batch, t, h, w, dim, "peg_bwd_weight");
  if (rc) return rc;
  if (dim % 32 == 0 && batch <= 65535)
    return launch_tiled<2>(x, dy, nullptr, nullptr, nullptr, nullptr, dw27, dbias, batch, t, h, w, dim, temporal, stream,
                           "peg_bwd_weight");
  const long long tokens = (long long)batch * t * h * w;
  int tokens_per_block = 128;
  long long blocks = (tokens + tokens_per_block - 1) / tokens_per_block;
  const long long cap = (long long)ctclip::sm_count() * 8;
  if (blocks > cap) {
    tokens_per_block = (int)((tokens + cap - 1) / cap);
    blocks = (tokens + tokens_per_block - 1) / tokens_per_block;
  }
  dim3 grid((unsigned)blocks, (unsigned)((dim / 4 + 127) / 128));
  peg_wgrad_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, dy, dw27, dbias, batch, PegGrid{t, h, w, temporal}, dim,
                                                           tokens_per_block);
  return ctclip::check_launch("peg_bwd_weight");
}
