// tcgen05 / TMEM / TMA GEMM for the CT-CLIP hot path (sm_100a only).
//
//   C[M,N] (op)= alpha * sum_k A(m,k) * B(n,k)  (+ bias[n]) (+ resid[m,n])
//
// bf16 operands, fp32 accumulation in tensor memory. Each operand may be stored K-major
// (rows = M or N, K contiguous) or MN-major (rows = K, M/N contiguous), so that forward
// (x W^T), dgrad (dy W) and wgrad (dy^T x) all run from the natural row-major tensors
// without transposes. Replaces the cuBLAS calls behind nn.Linear on the path
// (reference: CTPA_CLIP/ct_clip/ctvit.py:172, attention.py:48,51,119,120,125).
//
// Structure: persistent CTAs (one per SM), warp-specialised:
//   warp 0   : TMA producer (one elected lane)         smem ring: full/empty mbarriers
//   warp 1   : tcgen05.mma issuer (one elected lane)   TMEM ring: tmem_full/tmem_empty
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue (two per TMEM lane quadrant, half of the tile's columns each), tcgen05.ld 32x32b (one
//               accumulator row per thread, next chunk's load in flight while this chunk is stored) -> global
// Tile 128 x BN x 64, SWIZZLE_128B operand tiles, 2 accumulator buffers (2*BN TMEM columns).
#include "ptx.cuh"
#include "ctclip_internal.h"
#include <cstdlib>

namespace {

using namespace ptx;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 384;   // warps 0-2: TMA / MMA / TMEM alloc, warp 3 idle (keeps the epilogue warps quadrant-aligned), warps 4-11: epilogue
constexpr int kEpiWarps = 8;     // two per TMEM lane quadrant, each owning half of the tile's columns

template <int BN, bool PAIR = false>
struct SmemLayout {
  static constexpr int kStages = (BN == 256 && !PAIR) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * BK * 2;   // a CTA of a pair stages half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStageOut = kStages * kStageBytes;             // epilogue warps x 32 rows x 128 B staging
  static constexpr int kBarOffset = kStageOut + kEpiWarps * 32 * 128;
  static constexpr int kTotal = kBarOffset + 256 /*barriers + tmem ptr*/ + 1024 /*alignment slack*/;
};

struct GemmKernelParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_total, kb_per_split;
  void* C;
  long long ldc;
  int c_is_f32;
  int atomic;
  const float* bias;
  const float* resid;
  long long ldr;
  float alpha;
  float4* top2;  // non-null: per (row, n-tile) best two (value, column) instead of storing C
  // GEGLU forward epilogue (template EPI = 1; N = 2 Nh, tiles are 128-column slabs of the x half and of the gate half):
  //   B rows [x | gate]; accumulator columns [0,128) = x slab, [128,256) = gate slab; C <- h = [x | gate], u <- x gelu(gate).
  //   No fields of its own (a larger struct cost the other instantiations registers): Nh = N / 2, u = top2, ld_u = ldr.
  // batched mode (BERT attention): Z = zh_n * zb_n independent problems, z = b * zh_n + h
  int zh_n, z_n;
  int a_hpos, b_hpos;          // 1: tensor-map coordinates are (c0, h, row, b); 2: (c0, row, h, b)
  long long c_stride_h, c_stride_b;
};

// ---------------------------------------------------------------- epilogue helpers
__device__ __forceinline__ void store_row_chunk(const GemmKernelParams& p, int row, int col0, float (&v)[32]) {
  if (row >= p.M || col0 >= p.N) return;
  const int ncols = min(32, p.N - col0);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
  }
  if (p.c_is_f32) {
    float* c = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
    if (p.atomic) {
      if (ncols == 32 && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red_add_v4(c + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        return;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) atomicAdd(c + j, v[j]);
      return;
    }
    const bool vec = (ncols == 32) && ((reinterpret_cast<uintptr_t>(c) & 15) == 0) &&
                     (p.resid == nullptr ||
                      ((reinterpret_cast<uintptr_t>(p.resid + (long long)row * p.ldr + col0) & 15) == 0));
    if (vec) {
      if (p.resid != nullptr) {
        const float4* r4 = reinterpret_cast<const float4*>(p.resid + (long long)row * p.ldr + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 r = r4[j];
          v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      }
      float4* c4 = reinterpret_cast<float4*>(c);
#pragma unroll
      for (int j = 0; j < 8; ++j) c4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) {
          float x = v[j];
          if (p.resid != nullptr) x += p.resid[(long long)row * p.ldr + col0 + j];
          c[j] = x;
        }
    }
  } else {
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)row * p.ldc + col0;
    const bool vec = (ncols == 32) && ((reinterpret_cast<uintptr_t>(c) & 15) == 0);
    if (vec) {
      uint4* c4 = reinterpret_cast<uint4*>(c);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
        o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
        o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
        o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
        c4[j] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) c[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// Coalesced epilogue for a full, aligned 32-column chunk: the warp's 32x32 accumulator block (one row per thread) is
// transposed through a swizzled shared-memory staging tile so that global loads/stores cover whole 128-byte lines
// (fp32: 8 lanes x 16 B per row, 4 rows per instruction; bf16: 4 lanes x 16 B per row, 8 rows per instruction).
__device__ __forceinline__ void store_chunk_coalesced(const GemmKernelParams& p, uint8_t* stg, int row0, int col0,
                                                      const float (&v)[32], int lane) {
  if (p.c_is_f32) {
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8)
      *reinterpret_cast<float4*>(stg + lane * 128 + ((c8 ^ (lane & 7)) << 4)) =
          make_float4(v[4 * c8], v[4 * c8 + 1], v[4 * c8 + 2], v[4 * c8 + 3]);
    __syncwarp();
    const int c8 = lane & 7;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr) bias4 = *reinterpret_cast<const float4*>(p.bias + col0 + c8 * 4);
    // all residual loads of the chunk are issued before the first store (resid may alias C: no reordering by the compiler)
    float4 rr[8];
    if (p.resid != nullptr && !p.atomic) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int grow = row0 + it * 4 + (lane >> 3);
        rr[it] = (grow < p.M) ? *reinterpret_cast<const float4*>(p.resid + (long long)grow * p.ldr + col0 + c8 * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int it = 0; it < 8; ++it) rr[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = it * 4 + (lane >> 3);
      const int grow = row0 + i;
      float4 x = *reinterpret_cast<const float4*>(stg + i * 128 + ((c8 ^ (i & 7)) << 4));
      x.x += bias4.x + rr[it].x; x.y += bias4.y + rr[it].y; x.z += bias4.z + rr[it].z; x.w += bias4.w + rr[it].w;
      if (grow < p.M) {
        float* c = reinterpret_cast<float*>(p.C) + (long long)grow * p.ldc + col0 + c8 * 4;
        if (p.atomic) {
          red_add_v4(c, x.x, x.y, x.z, x.w);
        } else {
          *reinterpret_cast<float4*>(c) = x;
        }
      }
    }
    __syncwarp();
  } else {
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      float b[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) b[e] = v[8 * c4 + e] + (p.bias != nullptr ? __ldg(p.bias + col0 + 8 * c4 + e) : 0.f);
      *reinterpret_cast<uint4*>(stg + lane * 64 + ((c4 ^ sw) << 4)) =
          make_uint4(pack_bf16(b[0], b[1]), pack_bf16(b[2], b[3]), pack_bf16(b[4], b[5]), pack_bf16(b[6], b[7]));
    }
    __syncwarp();
    const int c4 = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int i = it * 8 + (lane >> 2);
      const int grow = row0 + i;
      const uint4 x = *reinterpret_cast<const uint4*>(stg + i * 64 + ((c4 ^ ((i >> 1) & 3)) << 4));
      if (grow < p.M)
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)grow * p.ldc + col0 + c4 * 8) = x;
    }
    __syncwarp();
  }
}

// Lean epilogues for the hot cases (alpha == 1, 16-byte aligned rows, whole 32-column chunks). The warp's accumulator
// block goes through a swizzled 4 KB staging tile so that every global access is a 128-bit piece of a full 128-byte line.
// bf16, no bias: 64 columns (two TMEM chunks) per round -> each output row receives one 128-byte line.
__device__ __forceinline__ void epi_store_bf16x64(const GemmKernelParams& p, uint8_t* stg, int row0, int col0,
                                                  const uint32_t (&a)[32], const uint32_t (&b)[32], int lane) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint32_t* s = g < 4 ? a : b;
    const int o = (g & 3) * 8;
    const uint4 v = make_uint4(pack_bf16(__uint_as_float(s[o + 0]), __uint_as_float(s[o + 1])),
                               pack_bf16(__uint_as_float(s[o + 2]), __uint_as_float(s[o + 3])),
                               pack_bf16(__uint_as_float(s[o + 4]), __uint_as_float(s[o + 5])),
                               pack_bf16(__uint_as_float(s[o + 6]), __uint_as_float(s[o + 7])));
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) = v;
  }
  __syncwarp();
  const int g = lane & 7;
  __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(p.C) + col0 + g * 8;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int i = it * 4 + (lane >> 3);
    const uint4 x = *reinterpret_cast<const uint4*>(stg + i * 128 + ((g ^ (i & 7)) << 4));
    if (row0 + i < p.M) *reinterpret_cast<uint4*>(cb + (long long)(row0 + i) * p.ldc) = x;
  }
  __syncwarp();
}
// fp32 (+ bias) (+ residual, which may alias C): 32 columns per round
template <bool RESID, bool BIAS>
__device__ __forceinline__ void epi_store_f32x32(const GemmKernelParams& p, uint8_t* stg, int row0, int col0,
                                                 const uint32_t (&a)[32], int lane) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) =
        make_uint4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]);
  __syncwarp();
  const int g = lane & 7;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (BIAS) bias4 = *reinterpret_cast<const float4*>(p.bias + col0 + g * 4);
  float* cb = reinterpret_cast<float*>(p.C) + col0 + g * 4;
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {   // residual loads of four row groups are batched ahead of their stores (resid may alias C)
    float4 rr[4];
    if (RESID) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int grow = row0 + (hb * 4 + it) * 4 + (lane >> 3);
        rr[it] = (grow < p.M) ? *reinterpret_cast<const float4*>(p.resid + (long long)grow * p.ldr + col0 + g * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int i = (hb * 4 + it) * 4 + (lane >> 3);
      float4 x = *reinterpret_cast<const float4*>(stg + i * 128 + ((g ^ (i & 7)) << 4));
      if (BIAS) { x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w; }
      if (RESID) { x.x += rr[it].x; x.y += rr[it].y; x.z += rr[it].z; x.w += rr[it].w; }
      if (row0 + i < p.M) *reinterpret_cast<float4*>(cb + (long long)(row0 + i) * p.ldc) = x;
    }
  }
  __syncwarp();
}

// GEGLU forward (attention.py:39-42): this warp's 32 rows x 32 columns of the x slab (xr) and of the gate slab (gr).
// h = [x | gate] and u = x * gelu(gate) leave as bf16 through the staging tile (row = [x 64 B | gate 64 B], 16-byte chunks
// XOR-swizzled with the row; 4 lanes x 16 B per row and array, 8 rows per instruction). u is computed from the ROUNDED x and
// gate, i.e. exactly what geglu_fwd_kernel computes from the stored h.
__device__ __forceinline__ void epi_geglu_fwd(const GemmKernelParams& p, uint8_t* stg, int row0, int col0,
                                              const uint32_t (&xr)[32], const uint32_t (&gr)[32], int lane) {
  const int nh = p.N >> 1;
  __nv_bfloat16* u_out = reinterpret_cast<__nv_bfloat16*>(p.top2);
  uint32_t ub[16];
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    uint32_t xb[4], gb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xb[k] = pack_bf16(__uint_as_float(xr[8 * c4 + 2 * k]), __uint_as_float(xr[8 * c4 + 2 * k + 1]));
      gb[k] = pack_bf16(__uint_as_float(gr[8 * c4 + 2 * k]), __uint_as_float(gr[8 * c4 + 2 * k + 1]));
      ub[4 * c4 + k] = pack_bf16(bf16_lo(xb[k]) * gelu_fast(bf16_lo(gb[k])), bf16_hi(xb[k]) * gelu_fast(bf16_hi(gb[k])));
    }
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((c4 ^ (lane & 7)) << 4)) = make_uint4(xb[0], xb[1], xb[2], xb[3]);
    *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + c4) ^ (lane & 7)) << 4)) = make_uint4(gb[0], gb[1], gb[2], gb[3]);
  }
  __syncwarp();
  const int c4 = lane & 3;
  const bool col_ok = col0 + c4 * 8 < nh;     // Nh is a multiple of 8: whole 16-byte pieces
  __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(p.C) + col0 + c4 * 8;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = it * 8 + (lane >> 2);
    const uint4 xv = *reinterpret_cast<const uint4*>(stg + i * 128 + ((c4 ^ (i & 7)) << 4));
    const uint4 gv = *reinterpret_cast<const uint4*>(stg + i * 128 + (((4 + c4) ^ (i & 7)) << 4));
    if (row0 + i < p.M && col_ok) {
      *reinterpret_cast<uint4*>(hb + (long long)(row0 + i) * p.ldc) = xv;
      *reinterpret_cast<uint4*>(hb + (long long)(row0 + i) * p.ldc + nh) = gv;
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) =
        make_uint4(ub[4 * q], ub[4 * q + 1], ub[4 * q + 2], ub[4 * q + 3]);
  __syncwarp();
  __nv_bfloat16* ubp = u_out + col0 + c4 * 8;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = it * 8 + (lane >> 2);
    const uint4 uv = *reinterpret_cast<const uint4*>(stg + i * 128 + ((c4 ^ (i & 7)) << 4));
    if (row0 + i < p.M && col_ok) *reinterpret_cast<uint4*>(ubp + (long long)(row0 + i) * p.ldr) = uv;
  }
  __syncwarp();
}

// fp32 + residual with the residual tile of the chunk ALREADY in registers (EPI = 3: loaded one chunk ahead by the caller, so
// that the global-load latency overlaps the previous chunk's transposition and stores instead of sitting on the critical path
// of every chunk). rr[it] = residual float4 of row (row0 + it * 4 + (lane >> 3)), columns col0 + 4 * (lane & 7) ...
__device__ __forceinline__ void epi_load_resid(const GemmKernelParams& p, int row0, int col0, int lane, float4 (&rr)[8]) {
  const float* rb = p.resid + col0 + (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int grow = row0 + it * 4 + (lane >> 3);
    rr[it] = (grow < p.M) ? *reinterpret_cast<const float4*>(rb + (long long)grow * p.ldr) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void epi_store_f32x32_rr(const GemmKernelParams& p, uint8_t* stg, int row0, int col0,
                                                    const uint32_t (&a)[32], const float4 (&rr)[8], int lane) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) =
        make_uint4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]);
  __syncwarp();
  const int g = lane & 7;
  float* cb = reinterpret_cast<float*>(p.C) + col0 + g * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int i = it * 4 + (lane >> 3);
    float4 x = *reinterpret_cast<const float4*>(stg + i * 128 + ((g ^ (i & 7)) << 4));
    x.x += rr[it].x; x.y += rr[it].y; x.z += rr[it].z; x.w += rr[it].w;
    if (row0 + i < p.M) *reinterpret_cast<float4*>(cb + (long long)(row0 + i) * p.ldc) = x;
  }
  __syncwarp();
}

// ---------------------------------------------------------------- kernel
// EPI = 1: GEGLU forward epilogue (a separate instantiation: its code must not cost the other epilogues registers — adding it as
// a run-time branch pushed their spills from < 100 to ~900 bytes and the fp32 + residual GEMMs lost 15 %).
// EPI = 3: fp32 + residual with the residual rows held one chunk ahead in registers (whole-chunk N only; the residual may not
//          alias C rows that a LATER chunk of the same warp reads... it never does: a chunk's residual is read before any store
//          of that chunk, and chunks touch disjoint columns).
// PAIR: two CTAs of a cluster (same TPC) compute a 256 x BN tile with tcgen05.mma.cta_group::2 — CTA r owns rows
// [128 r, 128 r + 128) of A and of the accumulator and stages HALF of the B tile (rows [BN/2 r, ...)), which cuts the
// shared-memory fill per k-block from 48 KB to 32 KB (6 stages instead of 4). Only the leader issues MMAs; both CTAs' TMA
// loads are counted on the leader's full barriers, MMA completion is multicast to both CTAs' empty / accumulator-full
// barriers, and both epilogues release the accumulator on the leader's barrier.
template <int BN, bool A_MN, bool B_MN, bool PAIR, int EPI = 0>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmKernelParams p) {
  using L = SmemLayout<BN, PAIR>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work units: (problem | split, m unit, n tile), n fastest. In PAIR mode an m unit is two m tiles (one per CTA) and the
  // two CTAs of a cluster walk the same unit sequence.
  const int cta_rank = PAIR ? (int)cluster_ctarank() : 0;
  const int m_units = PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int num_work = m_units * p.n_tiles * p.splits * (p.z_n > 1 ? p.z_n : 1);
  const int w_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_m = [&](int w) { const int mu = (w / p.n_tiles) % m_units; return PAIR ? 2 * mu + cta_rank : mu; };
  auto tile_zs = [&](int w) { return w / (p.n_tiles * m_units); };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full + a, 1);
      mbar_init(tmem_empty + a, PAIR ? 2 * kEpiWarps : kEpiWarps);   // pair: both CTAs' epilogues release on the leader
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_ptr_smem, 2 * BN);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_smem, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // barriers of BOTH CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = w_first; w < num_work; w += w_stride) {
        const int n_t = w % p.n_tiles;
        const int m_t = tile_m(w);
        const int zs = tile_zs(w);                        // split index, or problem index in batched mode
        const int sp = p.z_n > 1 ? 0 : zs;
        const int zh = p.z_n > 1 ? zs % p.zh_n : 0, zb = p.z_n > 1 ? zs / p.zh_n : 0;
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        // operand tiles: coordinates (contiguous, row) plus the (head, batch) pair of the problem
        auto load = [&](uint8_t* dst, const CUtensorMap* m, int hpos, int c0, int crow) {
          if constexpr (PAIR) {   // bytes are counted on the leader's full barrier
            if (hpos == 1) tma_load_4d_pair(dst, m, full_bar + stage, c0, zh, crow, zb);
            else tma_load_4d_pair(dst, m, full_bar + stage, c0, crow, zh, zb);
          } else {
            if (hpos == 1) tma_load_4d(dst, m, full_bar + stage, c0, zh, crow, zb);
            else tma_load_4d(dst, m, full_bar + stage, c0, crow, zh, zb);
          }
        };
        constexpr int kBRows = PAIR ? BN / 2 : BN;             // B rows this CTA stages
        // GEGLU forward: the tile's B rows are a 128-row slab of the x half and the matching slab of the gate half
        const int b_row0 = EPI == 1 ? n_t * 128 + (PAIR && cta_rank ? (p.N >> 1) : 0) : n_t * BN + (PAIR ? cta_rank * kBRows : 0);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if constexpr (PAIR) {
            if (cta_rank == 0) mbar_expect_tx(full_bar + stage, 2 * L::kStageBytes);   // leader: both CTAs' bytes
          } else {
            mbar_expect_tx(full_bar + stage, L::kStageBytes);
          }
          if constexpr (!A_MN) {
            load(sa, &tmap_a, p.a_hpos, kb * BK, m_t * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) load(sa + j * (BK * 128), &tmap_a, p.a_hpos, m_t * BM + j * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            load(sb, &tmap_b, p.b_hpos, kb * BK, b_row0);
            if constexpr (EPI == 1 && !PAIR) load(sb + 128 * 128, &tmap_b, p.b_hpos, kb * BK, (p.N >> 1) + b_row0);   // gate slab
          } else {
#pragma unroll
            for (int j = 0; j < kBRows / 64; ++j) load(sb + j * (BK * 128), &tmap_b, p.b_hpos, b_row0 + j * 64, kb * BK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: the leader CTA only) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = w_first; w < num_work; w += w_stride) {
        const int sp = p.z_n > 1 ? 0 : tile_zs(w);
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 8-row groups are 1024 B apart (SBO); a 16-element k step is +32 B inside the swizzle atom.
            // MN-major: 64-element MN atoms are BK*128 B apart (LBO), 8-row k groups 1024 B apart (SBO);
            //           a 16-row k step is +2048 B.
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, BK * 128, 1024)
                                     : make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, BK * 128, 1024)
                                     : make_smem_desc_sw128(sb + k * 32, 16, 1024);
            if constexpr (PAIR) mma_f16_ss_pair(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else mma_f16_ss(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (PAIR) mma_commit_pair(empty_bar + stage); else mma_commit(empty_bar + stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator ready for the epilogue(s)
        if constexpr (PAIR) mma_commit_pair(tmem_full + acc); else mma_commit(tmem_full + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp & 3;          // TMEM lane quadrant this warp may read
    const int ch = (warp - 4) >> 2;   // column half of the tile
    constexpr int kHalf = BN / 2, kChunks = kHalf / 32;
    uint8_t* stg = smem + L::kStageOut + (warp - 4) * (32 * 128);
    // 16-byte alignment of every row start of C / resid / bias -> the coalesced 128-bit epilogue is legal
    const bool aligned_out =
        ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.ldc * (p.c_is_f32 ? 4 : 2)) % 16 == 0) &&
        (p.resid == nullptr || (((reinterpret_cast<uintptr_t>(p.resid) & 15) == 0) && ((p.ldr * 4) % 16 == 0))) &&
        (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) &&
        (p.z_n <= 1 || (((p.c_stride_h | p.c_stride_b) * (p.c_is_f32 ? 4 : 2)) % 16 == 0));
    // 0: generic; 1: bf16 plain; 2..5: fp32 (+resid) (+bias)
    const int fast_mode = (!aligned_out || p.atomic || p.alpha != 1.f) ? 0
                          : (!p.c_is_f32 ? (p.bias == nullptr ? 1 : 0)
                                         : 2 + (p.resid != nullptr ? 1 : 0) + (p.bias != nullptr ? 2 : 0));
    int acc = 0;
    uint32_t acc_phase = 0;
    GemmKernelParams pz = p;   // per-problem view: C shifted to the (head, batch) slice in batched mode
    for (int w = w_first; w < num_work; w += w_stride) {
      const int n_t = w % p.n_tiles;
      const int m_t = tile_m(w);
      if (p.z_n > 1) {
        const int zs = tile_zs(w);
        const long long off = (long long)(zs % p.zh_n) * p.c_stride_h + (long long)(zs / p.zh_n) * p.c_stride_b;
        pz.C = p.c_is_f32 ? static_cast<void*>(reinterpret_cast<float*>(p.C) + off)
                          : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(p.C) + off);
      }
      if (EPI != 3 && (fast_mode == 3 || fast_mode == 5)) {
        // The residual rows of the NEXT work unit start their way from HBM into L2 now (no registers held): one tile
        // epilogue later the loads below are L2 hits, so the few loads a warp keeps in flight no longer bound the bandwidth
        auto prefetch_resid = [&](int wq) {
          const int prow = tile_m(wq) * BM + ew * 32 + lane;
          const int pcol = (wq % p.n_tiles) * BN + ch * kHalf;
          if (prow < p.M) {
            const char* base = reinterpret_cast<const char*>(p.resid + (long long)prow * p.ldr + pcol);
#pragma unroll
            for (int k = 0; k < kHalf / 32; ++k)
              if (pcol + k * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + k * 128));
          }
        };
        if (w == w_first) prefetch_resid(w);
        if (w + w_stride < num_work) prefetch_resid(w + w_stride);
      }
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      const int row = m_t * BM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
      if constexpr (EPI == 3) {
        const int cbase = n_t * BN + ch * kHalf;
        const int row0 = m_t * BM + ew * 32;
        uint32_t r[32];
        float4 rr0[8];
        // (chunk 0 was requested before the accumulator wait)
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          if (c == 0) epi_load_resid(pz, row0, cbase, lane, rr0);
          tmem_ld_32x32(taddr + ch * kHalf + c * 32, r);
          float4 rn[8];
          if (c + 1 < kChunks) epi_load_resid(pz, row0, cbase + (c + 1) * 32, lane, rn);
          tmem_wait_ld();
          epi_store_f32x32_rr(pz, stg, row0, cbase + c * 32, r, rr0, lane);
          if (c + 1 < kChunks) {
#pragma unroll
            for (int it = 0; it < 8; ++it) rr0[it] = rn[it];
          }
        }
      } else if constexpr (EPI == 1) {
        // GEGLU forward: this warp owns slab columns [64 ch, 64 ch + 64) of BOTH slabs (x at TMEM column lc, gate at 128 + lc)
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int lc = ch * 64 + c * 32;
          const int gc = n_t * 128 + lc;           // column of x inside h; the gate column is Nh + gc
          if (gc >= (p.N >> 1)) break;             // warp-uniform
          uint32_t xr[32], gr[32];
          tmem_ld_32x32(taddr + lc, xr);
          tmem_ld_32x32(taddr + 128 + lc, gr);
          tmem_wait_ld();
          epi_geglu_fwd(pz, stg, m_t * BM + ew * 32, gc, xr, gr, lane);
        }
      } else if (p.top2 != nullptr) {
        // VQ assignment epilogue: running best-two over this tile's columns (ties keep the lower column). The two warps of
        // a lane quadrant scan one half of the columns each; the upper half hands its pair to the lower one through the
        // staging tile (named barrier of the two warps), which merges and writes one float4 per row.
        float v1 = -INFINITY, v2 = -INFINITY;
        int i1 = -1, i2 = -1;
#pragma unroll 1
        for (int c = ch * kHalf; c < (ch + 1) * kHalf; c += 32) {
          if (n_t * BN + c >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = n_t * BN + c + j;
            const float x = __uint_as_float(r[j]);
            if (col < p.N) {
              if (x > v1) { v2 = v1; i2 = i1; v1 = x; i1 = col; }
              else if (x > v2) { v2 = x; i2 = col; }
            }
          }
        }
        float4* xch = reinterpret_cast<float4*>(smem + L::kStageOut + (4 + ew) * (32 * 128));   // upper-half warp's tile
        if (ch == 1) xch[lane] = make_float4(v1, __int_as_float(i1), v2, __int_as_float(i2));
        asm volatile("bar.sync %0, 64;" ::"r"(1 + ew) : "memory");
        if (ch == 0) {
          const float4 u = xch[lane];
          const float b1 = u.x, b2 = u.z;
          const int bi1 = __float_as_int(u.y), bi2 = __float_as_int(u.w);
          float t1, t2; int j1, j2;
          if (b1 > v1) {          // upper columns win only when strictly larger
            t1 = b1; j1 = bi1;
            if (b2 > v1) { t2 = b2; j2 = bi2; } else { t2 = v1; j2 = i1; }
          } else {
            t1 = v1; j1 = i1;
            if (b1 > v2) { t2 = b1; j2 = bi1; } else { t2 = v2; j2 = i2; }
          }
          if (row < p.M)
            p.top2[(long long)row * p.n_tiles + n_t] = make_float4(t1, __int_as_float(j1), t2, __int_as_float(j2));
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + ew) : "memory");   // the pair's tile is free for the next work unit
      } else if (fast_mode == 1 && n_t * BN + (ch + 1) * kHalf <= p.N) {
        // bf16 plain: 64 columns per round; the next round's TMEM loads fly while this round is stored
        const int cbase = n_t * BN + ch * kHalf;
        uint32_t r[2][32];
        tmem_ld_32x32(taddr + ch * kHalf, r[0]);
        tmem_ld_32x32(taddr + ch * kHalf + 32, r[1]);
#pragma unroll
        for (int c = 0; c < kChunks; c += 2) {
          tmem_wait_ld();
          uint32_t a[32], b[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { a[j] = r[0][j]; b[j] = r[1][j]; }
          if (c + 2 < kChunks) {
            tmem_ld_32x32(taddr + ch * kHalf + (c + 2) * 32, r[0]);
            tmem_ld_32x32(taddr + ch * kHalf + (c + 3) * 32, r[1]);
          }
          epi_store_bf16x64(pz, stg, m_t * BM + ew * 32, cbase + c * 32, a, b, lane);
        }
      } else if (fast_mode >= 2 && n_t * BN + (ch + 1) * kHalf <= p.N) {
        const int cbase = n_t * BN + ch * kHalf;
        uint32_t r[2][32];
        tmem_ld_32x32(taddr + ch * kHalf, r[0]);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          tmem_wait_ld();
          if (c + 1 < kChunks) tmem_ld_32x32(taddr + ch * kHalf + (c + 1) * 32, r[(c + 1) & 1]);
          const int row0 = m_t * BM + ew * 32, col0 = cbase + c * 32;
          if (fast_mode == 2) epi_store_f32x32<false, false>(pz, stg, row0, col0, r[c & 1], lane);
          else if (fast_mode == 3) epi_store_f32x32<true, false>(pz, stg, row0, col0, r[c & 1], lane);
          else if (fast_mode == 4) epi_store_f32x32<false, true>(pz, stg, row0, col0, r[c & 1], lane);
          else epi_store_f32x32<true, true>(pz, stg, row0, col0, r[c & 1], lane);
        }
      } else {
        const int cbase = n_t * BN + ch * kHalf;
        uint32_t r[2][32];
        if (cbase < p.N) tmem_ld_32x32(taddr + ch * kHalf, r[0]);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int col0 = cbase + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          tmem_wait_ld();
          if (c + 1 < kChunks && col0 + 32 < p.N) tmem_ld_32x32(taddr + ch * kHalf + (c + 1) * 32, r[(c + 1) & 1]);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[c & 1][j]) * p.alpha;
          if (aligned_out && col0 + 32 <= p.N)
            store_chunk_coalesced(pz, stg, m_t * BM + ew * 32, col0, v, lane);
          else
            store_row_chunk(pz, row, col0, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_leader(tmem_empty + acc); else mbar_arrive(tmem_empty + acc);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // pair: the peer's smem / TMEM stay alive until both are done
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------- host side
// 4-D operand map: (contiguous dim, rows) of one problem plus the (head, batch) dimensions of batched mode; the two
// middle dimensions are ordered by stride. K-major: the tensor is [rows_mn][k] (k contiguous), box {64, box_mn};
// MN-major: [k][rows_mn] (mn contiguous), box {64, 64}. Extents are per problem, so every edge is TMA zero fill.
int encode_operand_map(CUtensorMap* map, int* hpos, const void* ptr, bool mn_major, long long rows_mn, long long k,
                       long long ld_elems, int box_mn, int zh_n, int zb_n, long long stride_h, long long stride_b) {
  const cuuint64_t inner = mn_major ? rows_mn : k, rows = mn_major ? k : rows_mn;
  const cuuint32_t box_rows = mn_major ? BK : box_mn;
  if (zh_n <= 1) stride_h = (long long)rows * ld_elems;        // extent-1 dimensions: any legal stride
  if (zb_n <= 1) stride_b = (stride_h > (long long)rows * ld_elems ? stride_h : (long long)rows * ld_elems) *
                            (zh_n > 1 ? zh_n : 1);
  const bool h_first = zh_n > 1 && stride_h < ld_elems;        // e.g. heads interleaved inside a token row
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  dims[0] = inner; box[0] = 64;
  if (h_first) {
    dims[1] = zh_n > 1 ? zh_n : 1; strides[0] = (cuuint64_t)stride_h * 2; box[1] = 1;
    dims[2] = rows; strides[1] = (cuuint64_t)ld_elems * 2; box[2] = box_rows;
  } else {
    dims[1] = rows; strides[0] = (cuuint64_t)ld_elems * 2; box[1] = box_rows;
    dims[2] = zh_n > 1 ? zh_n : 1; strides[1] = (cuuint64_t)stride_h * 2; box[2] = 1;
  }
  dims[3] = zb_n > 1 ? zb_n : 1; strides[2] = (cuuint64_t)stride_b * 2; box[3] = 1;
  *hpos = h_first ? 1 : 2;
  return ctclip::encode_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                             estr, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN, bool A_MN, bool B_MN, bool PAIR, int EPI = 0>
int launch(const ctclip_gemm_desc* d, const GemmKernelParams& kp, const CUtensorMap& ta, const CUtensorMap& tb,
           int grid, cudaStream_t stream) {
  using L = SmemLayout<BN, PAIR>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, PAIR, EPI>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  if constexpr (PAIR) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, kp);
    if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "gemm (CTA pair) launch: %s", cudaGetErrorString(e));
  } else {
    kern<<<grid, kThreads, L::kTotal, stream>>>(ta, tb, kp);
  }
  (void)d;
  return ctclip::check_launch("gemm launch");
}

// CTA-pair mode is used for the wide (BN = 256) tiles unless CTCLIP_GEMM_PAIR=0
bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CTCLIP_GEMM_PAIR");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace

extern "C" int ctclip_gemm_bf16(const ctclip_gemm_desc* d, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (d == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "gemm: null descriptor");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return ctclip::fail(CTCLIP_E_SHAPE, "gemm: M,N,K must be positive");
  if ((d->lda % 8) || (d->ldb % 8)) return ctclip::fail(CTCLIP_E_ALIGN, "gemm: lda/ldb must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(d->A) & 15) || (reinterpret_cast<uintptr_t>(d->B) & 15))
    return ctclip::fail(CTCLIP_E_ALIGN, "gemm: A/B must be 16-byte aligned");
  if (d->A == nullptr || d->B == nullptr || (d->C == nullptr && d->top2_out == nullptr)) return ctclip::fail(CTCLIP_E_SHAPE, "gemm: null pointer");
  int rc = ctclip::require_sm100();
  if (rc) return rc;

  const bool geglu_fwd = d->geglu_u != nullptr;
  if (d->geglu_h != nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "gemm: the GEGLU backward epilogue (geglu_h) is not implemented");
  if (geglu_fwd) {
    if ((d->N % 16) || d->a_mn_major || d->b_mn_major || d->c_is_f32 || d->bias || d->resid || d->atomic || d->top2_out ||
        d->splits > 1 || d->alpha != 1.f || d->batch_h > 1 || d->batch_b > 1 || d->C == nullptr)
      return ctclip::fail(CTCLIP_E_SHAPE, "gemm: the GEGLU forward epilogue needs N = 2 Nh with Nh %% 8 == 0, K-major A and B, "
                                          "bf16 C and no bias / resid / atomic / top2 / split-K / alpha / batch");
    if ((d->ldc % 8) || (d->ld_u % 8) || (reinterpret_cast<uintptr_t>(d->C) & 15) || (reinterpret_cast<uintptr_t>(d->geglu_u) & 15))
      return ctclip::fail(CTCLIP_E_ALIGN, "gemm: GEGLU outputs need 16-byte aligned rows");
  }
  const int BN = (d->N <= 128 && !geglu_fwd) ? 128 : 256;
  GemmKernelParams kp{};
  kp.M = d->M; kp.N = d->N; kp.K = d->K;
  kp.m_tiles = (d->M + BM - 1) / BM;
  kp.n_tiles = geglu_fwd ? (d->N / 2 + 127) / 128 : (d->N + BN - 1) / BN;

  kp.kb_total = (d->K + BK - 1) / BK;
  int splits = d->splits;
  const int tiles = kp.m_tiles * kp.n_tiles;
  const int sms = ctclip::sm_count();
  if (splits <= 0) {  // auto: only when the caller allows atomic accumulation
    splits = 1;
    if (d->atomic) {
      // cost model in k-block units: the persistent CTAs take ceil(units / SMs) work units each; a unit is its share of the
      // k-blocks plus a fixed cost (pipeline fill + the vector-RED epilogue of a 128 x BN fp32 tile ~ 8 k-blocks of MMA time)
      int max_splits = kp.kb_total < 128 ? kp.kb_total : 128;
      const long long kFixed = 8;
      long long best_cost = -1;
      for (int sp = 1; sp <= max_splits; ++sp) {
        const long long units = (long long)tiles * sp;
        const long long waves = (units + sms - 1) / sms;
        const long long cost = waves * ((kp.kb_total + sp - 1) / sp + kFixed);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = sp; }
      }
    }
  }
  if (splits > 1 && !(d->atomic && d->c_is_f32))
    return ctclip::fail(CTCLIP_E_SHAPE, "gemm: split-K requires atomic fp32 output");
  if (d->atomic && !d->c_is_f32) return ctclip::fail(CTCLIP_E_SHAPE, "gemm: atomic output must be fp32");
  kp.kb_per_split = (kp.kb_total + splits - 1) / splits;
  kp.splits = (kp.kb_total + kp.kb_per_split - 1) / kp.kb_per_split;  // no empty splits
  kp.C = d->C; kp.ldc = d->ldc; kp.c_is_f32 = d->c_is_f32; kp.atomic = d->atomic;
  kp.bias = d->bias; kp.resid = d->resid; kp.ldr = d->ldr;
  kp.alpha = d->alpha;
  kp.top2 = reinterpret_cast<float4*>(d->top2_out);
  if (geglu_fwd) { kp.top2 = reinterpret_cast<float4*>(d->geglu_u); kp.ldr = d->ld_u; }   // EPI = 1 reads them as (u, ld_u)
  if (kp.top2 != nullptr && kp.splits > 1) return ctclip::fail(CTCLIP_E_SHAPE, "gemm: top2 epilogue cannot be split-K");
  if (kp.splits > 1 && (d->bias || d->resid))
    return ctclip::fail(CTCLIP_E_SHAPE, "gemm: bias/resid not supported with split-K");

  const int zh_n = d->batch_h > 1 ? d->batch_h : 1, zb_n = d->batch_b > 1 ? d->batch_b : 1;
  kp.zh_n = zh_n; kp.z_n = zh_n * zb_n;
  kp.c_stride_h = d->c_stride_h; kp.c_stride_b = d->c_stride_b;
  if (kp.z_n > 1) {
    if (kp.splits > 1 || d->atomic || d->resid || kp.top2)
      return ctclip::fail(CTCLIP_E_SHAPE, "gemm: batched mode supports neither split-K / atomic output nor resid / top2");
    if ((d->a_stride_h % 8) || (d->a_stride_b % 8) || (d->b_stride_h % 8) || (d->b_stride_b % 8))
      return ctclip::fail(CTCLIP_E_ALIGN, "gemm: batch strides must be multiples of 8 elements");
  }
  // CTA pairs pay off where the tensor pipe is the limit: deep enough K, enough tiles to fill every pair
  const bool pair = BN == 256 && pair_enabled() && kp.m_tiles >= 2 && d->K >= 512 &&
                    (long long)tiles * kp.splits * kp.z_n >= sms;
  CUtensorMap ta, tb;
  rc = encode_operand_map(&ta, &kp.a_hpos, d->A, d->a_mn_major != 0, d->M, d->K, d->lda, BM, zh_n, zb_n, d->a_stride_h,
                          d->a_stride_b);
  if (rc) return rc;
  rc = encode_operand_map(&tb, &kp.b_hpos, d->B, d->b_mn_major != 0, d->N, d->K, d->ldb, (pair || geglu_fwd) ? BN / 2 : BN, zh_n, zb_n,
                          d->b_stride_h, d->b_stride_b);
  if (rc) return rc;

  // fp32 + residual, whole 256-column tiles, K-major operands: the register-prefetching instantiation (CTCLIP_GEMM_RESID_AHEAD=0: off)
  bool resid_ahead = false;
  {
    const char* e = getenv("CTCLIP_GEMM_RESID_AHEAD");
    const bool on = !(e != nullptr && e[0] == '0');
    resid_ahead = on && d->resid != nullptr && d->c_is_f32 && d->bias == nullptr && !d->atomic && d->alpha == 1.f && kp.z_n == 1 &&
                  kp.top2 == nullptr && !geglu_fwd && BN == 256 && d->N % 256 == 0 && !d->a_mn_major && !d->b_mn_major &&
                  (reinterpret_cast<uintptr_t>(d->C) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->resid) & 15) == 0 &&
                  (d->ldc % 4) == 0 && (d->ldr % 4) == 0;
  }

  if (pair) {
    const int units = ((kp.m_tiles + 1) / 2) * kp.n_tiles * kp.splits * kp.z_n;
    int clusters = sms / 2;
    if (clusters > units) clusters = units;
    const int grid = 2 * clusters;
    const int sel = (d->a_mn_major ? 2 : 0) | (d->b_mn_major ? 1 : 0);
    if (geglu_fwd) return launch<256, false, false, true, 1>(d, kp, ta, tb, grid, stream);
    if (resid_ahead) return launch<256, false, false, true, 3>(d, kp, ta, tb, grid, stream);
    switch (sel) {
      case 0: return launch<256, false, false, true>(d, kp, ta, tb, grid, stream);
      case 1: return launch<256, false, true, true>(d, kp, ta, tb, grid, stream);
      case 2: return launch<256, true, false, true>(d, kp, ta, tb, grid, stream);
      default: return launch<256, true, true, true>(d, kp, ta, tb, grid, stream);
    }
  }
  const int num_work = tiles * kp.splits * kp.z_n;
  const int grid = num_work < sms ? num_work : sms;
  const int sel = (BN == 256 ? 4 : 0) | (d->a_mn_major ? 2 : 0) | (d->b_mn_major ? 1 : 0);
  if (geglu_fwd) return launch<256, false, false, false, 1>(d, kp, ta, tb, grid, stream);
  if (resid_ahead) return launch<256, false, false, false, 3>(d, kp, ta, tb, grid, stream);
  switch (sel) {
    case 0: return launch<128, false, false, false>(d, kp, ta, tb, grid, stream);
    case 1: return launch<128, false, true, false>(d, kp, ta, tb, grid, stream);
    case 2: return launch<128, true, false, false>(d, kp, ta, tb, grid, stream);
    case 3: return launch<128, true, true, false>(d, kp, ta, tb, grid, stream);
    case 4: return launch<256, false, false, false>(d, kp, ta, tb, grid, stream);
    case 5: return launch<256, false, true, false>(d, kp, ta, tb, grid, stream);
    case 6: return launch<256, true, false, false>(d, kp, ta, tb, grid, stream);
    default: return launch<256, true, true, false>(d, kp, ta, tb, grid, stream);
  }
}

extern "C" int ctclip_gemm_tile_n(int N) { return (N <= 128) ? 128 : 256; }
