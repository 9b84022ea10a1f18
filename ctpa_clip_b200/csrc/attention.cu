// Cosine-sim attention of the CTViT transformer blocks on tcgen05 / TMEM (reference attention.py:127-181):
//   q = l2norm(q) * q_scale ; k = l2norm(k) * k_scale ; P = softmax(8 q k^T + bias) ; O = P v      (dim_head = 32)
//
// One CTA owns one (sequence block, head): its K^, V (and, in the backward, Q^, dO) live in shared memory as bf16
// 8x8 "core matrices" (SWIZZLE_NONE UMMA layout), written by the threads that normalise them, so the same bytes
// serve as K-major or MN-major operands. S = Q^ K^T and O = P V are tcgen05.mma with accumulators in tensor
// memory; tcgen05.ld 32x32b hands every thread one full query row, so the softmax needs no shuffles.
// No running max: |8 q^.k^| <= 8 max_d|qs_d ks_d| and the additive bias has a per-query-position maximum known
// from the (2h-1)(2w-1) relative-position table, so exp2(s - M_i) can never overflow and never fully underflows.
//
// Both factorised attentions share this kernel: "spatial" blocks are one frame of n = h*w tokens (bias from the
// continuous-position-bias table); "temporal" blocks pack floor(128/t) sequences of t tokens (gathered with stride
// h*w from the canonical (b,t,h,w,d) layout) and mask pairs from different sequences.
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {
using namespace ptx;

constexpr int DH = 32;            // head dim
constexpr int QT = 128;           // query rows per tile
constexpr int KC = 64;            // keys per chunk (forward)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kScale = 8.0f;    // attention.py:99

struct AttnParams {
  const __nv_bfloat16* q;   // [tokens][ldq]
  const __nv_bfloat16* kv;  // [tokens][ldkv]  k = cols [0,inner), v = cols [inner, 2 inner)
  __nv_bfloat16* o;         // [tokens][ldo]
  float* lse;               // [tokens][heads]  (log2 domain, includes the offset)
  const float* q_scale;     // [32]
  const float* k_scale;     // [32]
  const float* bias_table;  // [heads][(2h-1)*(2w-1)] or null
  const float* bias_rowmax; // [heads][h*w] or null
  int heads, inner, ldq, ldkv, ldo;
  int mode;                 // 0 spatial, 1 temporal
  int n;                    // tokens per sequence
  int nst;                  // rows a sequence occupies inside a block (n, or n padded to 32 / 64 for packed short sequences)
  int ns;                   // sequences per block
  int num_seqs;             // total sequences
  int gh, gw, gt;           // token grid
  int r_pad;                // rows per block padded to a multiple of 64
  int num_blocks;           // sequence blocks per head; a CTA (fixed head) walks blocks blockIdx.x / heads, + gridDim.x / heads, ...
};

// Row pitch of the relative-position bias table inside shared memory: the smallest S >= 2 gw - 1 with S = gw (mod 32). The
// lanes of a warp are consecutive query positions p = qy * gw + qx, and they read table word A_i - B_j with
// A_i = (qy + gh - 1) S + qx + gw - 1: with this pitch A_i = p + const (mod 32), so the 32 lanes always hit 32 different banks
// (with the natural pitch 2 gw - 1 = 47 lanes of neighbouring grid rows collide: 35 % of the backward kernel's shared-memory
// wavefronts were bank conflicts of the bias loads and of the dbias read-modify-writes).
__host__ __device__ __forceinline__ int tab_pitch(int gw) { return gw + 32 * ((gw - 1 + 31) / 32); }
__host__ __device__ __forceinline__ int tab_words(int gh, int gw) { return (2 * gh - 2) * tab_pitch(gw) + 2 * gw - 1; }

__device__ __forceinline__ long long row_token(const AttnParams& p, int blk, int r, bool& valid) {
  const int sl = r / p.nst, pos = r - sl * p.nst;
  const long long seq = (long long)blk * p.ns + sl;
  valid = (sl < p.ns) && (pos < p.n) && (seq < p.num_seqs);
  if (!valid) return 0;
  if (p.mode == 0) return seq * p.n + pos;
  const int hw = p.gh * p.gw;
  const long long b = seq / hw;
  const int ph = (int)(seq - b * hw);
  return (b * p.gt + pos) * hw + ph;
}

__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// The same descriptor as two 32-bit words: only the low word depends on the address, so the MMA-issuing thread derives
// every descriptor of a step from a few base words with one 32-bit add each (byte offset >> 4; shared memory addresses
// stay below 2^18, no carry out of the 14-bit field) instead of rebuilding 64-bit descriptors per instruction.
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) {
  return ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void mma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// byte offset of the 16-byte group (row r, column group c8) in a core-matrix tile with `cols` columns
__device__ __forceinline__ uint32_t cm_off(int r, int c8, int cols) {
  return (uint32_t)((r >> 3) * (cols >> 3) * 128 + c8 * 128 + (r & 7) * 16);
}

// loads one 32-wide bf16 head slice, returns fp32 values and the inverse l2 norm (F.normalize eps 1e-12)
__device__ __forceinline__ float load_head_row(const __nv_bfloat16* src, float (&v)[DH]) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = s4[j];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[8 * j + 2 * k] = bf16_lo(w[k]);
      v[8 * j + 2 * k + 1] = bf16_hi(w[k]);
    }
  }
#pragma unroll
  for (int j = 0; j < DH; ++j) ss = fmaf(v[j], v[j], ss);
  return 1.f / fmaxf(sqrtf(ss), 1e-12f);
}

// unpack a raw 32-wide bf16 head row held in registers, return fp32 values and the inverse l2 norm
__device__ __forceinline__ float unpack_head_row(const uint4 (&raw)[4], float (&v)[DH]) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t w[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[8 * j + 2 * k] = bf16_lo(w[k]);
      v[8 * j + 2 * k + 1] = bf16_hi(w[k]);
    }
  }
#pragma unroll
  for (int j = 0; j < DH; ++j) ss = fmaf(v[j], v[j], ss);
  return 1.f / fmaxf(sqrtf(ss), 1e-12f);
}

__device__ __forceinline__ void store_row_cm(uint8_t* tile, int r, const float (&v)[DH]) {
#pragma unroll
  for (int c8 = 0; c8 < 4; ++c8) {
    uint4 u;
    u.x = pack_bf16(v[8 * c8 + 0], v[8 * c8 + 1]);
    u.y = pack_bf16(v[8 * c8 + 2], v[8 * c8 + 3]);
    u.z = pack_bf16(v[8 * c8 + 4], v[8 * c8 + 5]);
    u.w = pack_bf16(v[8 * c8 + 6], v[8 * c8 + 7]);
    *reinterpret_cast<uint4*>(tile + cm_off(r, c8, DH)) = u;
  }
}
__device__ __forceinline__ void store_zero_row_cm(uint8_t* tile, int r) {
#pragma unroll
  for (int c8 = 0; c8 < 4; ++c8) *reinterpret_cast<uint4*>(tile + cm_off(r, c8, DH)) = make_uint4(0, 0, 0, 0);
}

// =============================================================================================== forward
// grid.x = num_blocks * heads (head fastest), 256 threads: thread t owns query row (t & 127) of the current tile and the
// column half (t >> 7) of every 64-key chunk. Per element: one broadcast LDS (key offset), one table LDS, 2 FADD, 1 EX2.
// The additive bias index is A_i - B_j with A_i = (qy+gh-1)(2gw-1) + qx+gw-1 per query row and B_j = ky(2gw-1)+kx per key.
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int head = blockIdx.x % p.heads;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowt = tid & 127, half = tid >> 7;
  const int R = p.ns * p.nst;
  const int nchunks = (R + KC - 1) / KC;
  const int ntiles = (R + QT - 1) / QT;
  const bool has_bias = p.bias_table != nullptr;

  const int kv_rows = nchunks * KC;          // keys padded to whole chunks
  uint8_t* sK = smem;                        // kv_rows x 32 bf16
  uint8_t* sV = sK + kv_rows * DH * 2;       // kv_rows x 32
  uint8_t* sQ = sV + kv_rows * DH * 2;       // 128 x 32
  uint8_t* sP = sQ + QT * DH * 2;            // 128 x 64
  int* sB = reinterpret_cast<int*>(sP + QT * KC * 2);   // [r_pad] key offsets B_j (16-byte aligned)
  float* sL = reinterpret_cast<float*>(sB + p.r_pad);   // [2][128] partial row sums of the two column halves
  float* sScale = sL + 256;                             // [64]: q_scale*8*log2e , k_scale
  float* sTab = sScale + 64;
  const int tab_n = has_bias ? (2 * p.gh - 1) * (2 * p.gw - 1) : 0;   // words of the table in global memory
  const int tab_s = has_bias ? tab_words(p.gh, p.gw) : 0;              // words of the padded copy in shared memory
  const int tpitch = tab_pitch(p.gw), tw = 2 * p.gw - 1;
  float* sRowMax = sTab + tab_s;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(sRowMax + (has_bias ? p.n : 0)) + 15) & ~uintptr_t(15));
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  if (tid == 0) {
    mbar_init(bars + 0, 1);  // S buffer 0 ready
    mbar_init(bars + 1, 1);  // S buffer 1 ready
    mbar_init(bars + 2, 1);  // P V of the previous chunk retired (P tile free, O updated)
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  if (tid < DH) {
    sScale[tid] = p.q_scale[tid] * (kScale * kLog2e);
    sScale[DH + tid] = p.k_scale[tid];
  }
  for (int i = tid; i < tab_n; i += blockDim.x)
    sTab[(i / tw) * tpitch + i % tw] = p.bias_table[(long long)head * tab_n + i] * kLog2e;
  if (has_bias)
    for (int i = tid; i < p.n; i += blockDim.x) sRowMax[i] = p.bias_rowmax[(long long)head * p.n + i] * kLog2e;
  for (int r = tid; r < p.r_pad; r += blockDim.x) {
    const int kp = r % p.n;
    sB[r] = -4 * ((kp / p.gw) * tpitch + (kp % p.gw));   // byte offset of key r inside the (padded) bias table
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  float cmax = 0.f;  // bound of |8 log2e q^.k^|
#pragma unroll
  for (int d = 0; d < DH; ++d) cmax = fmaxf(cmax, fabsf(sScale[d] * sScale[DH + d]));

  const uint32_t tmem = *tmem_ptr;
  const uint32_t tS = tmem;         // 2 x 64 columns
  const uint32_t tO = tmem + 128;   // 32 columns
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  constexpr uint32_t idesc_s = make_idesc_bf16(QT, KC, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(QT, DH, false, true);
  uint32_t ph_s0 = 0, ph_s1 = 0, ph_o = 0;

  // raw Q row of a tile (half 0 threads), loaded one tile ahead so that its global latency hides behind the key loop
  uint4 qraw[4];
  bool qvalid = false;
  long long qtok = 0;
  auto load_q_row = [&](int blk_, int tile_) {
    const int r_ = tile_ * QT + rowt;
    qvalid = false;
    qtok = (r_ < R) ? row_token(p, blk_, r_, qvalid) : 0;
    if (half == 0 && qvalid) {
      const uint4* src = reinterpret_cast<const uint4*>(p.q + qtok * p.ldq + head * DH);
#pragma unroll
      for (int j = 0; j < 4; ++j) qraw[j] = src[j];
    }
  };

  // persistent over the sequence blocks of this head: tables, barriers and TMEM are set up once per CTA
  for (int blk = blockIdx.x / p.heads; blk < p.num_blocks; blk += gridDim.x / p.heads) {
  load_q_row(blk, 0);   // in flight together with the K / V rows below
  // ---- K^ and V for the whole block (every MMA of the previous block has retired: its last tile waited bars + 2)
  for (int r = tid; r < kv_rows; r += blockDim.x) {
    bool valid = false;
    const long long tok = (r < R) ? row_token(p, blk, r, valid) : 0;
    if (valid) {
      float v[DH];
      const float inv = load_head_row(p.kv + tok * p.ldkv + head * DH, v);
#pragma unroll
      for (int d = 0; d < DH; ++d) v[d] *= inv * sScale[DH + d];
      store_row_cm(sK, r, v);
      const uint4* s4 = reinterpret_cast<const uint4*>(p.kv + tok * p.ldkv + p.inner + head * DH);
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) *reinterpret_cast<uint4*>(sV + cm_off(r, c8, DH)) = s4[c8];
    } else {
      store_zero_row_cm(sK, r);
      store_zero_row_cm(sV, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  for (int tile = 0; tile < ntiles; ++tile) {
    const int r = tile * QT + rowt;
    const bool valid = qvalid;
    const long long tok = qtok;
    if (half == 0) {
      if (valid) {
        float v[DH];
        const float inv = unpack_head_row(qraw, v);
#pragma unroll
        for (int d = 0; d < DH; ++d) v[d] *= inv * sScale[d];
        store_row_cm(sQ, rowt, v);
      } else {
        store_zero_row_cm(sQ, rowt);
      }
    }
    if (tile + 1 < ntiles) load_q_row(blk, tile + 1);   // next tile's row: consumed after this tile's key loop
    const int my_seq = r / p.nst;
    const int my_pos = r - my_seq * p.nst;
    const int key_lo = my_seq * p.nst, key_hi = min(R, key_lo + p.n);   // keys of this row's own sequence
    const int a_i = (my_pos / p.gw + p.gh - 1) * tpitch + (my_pos % p.gw) + p.gw - 1;
    const float m_i = cmax + ((has_bias && valid) ? sRowMax[my_pos] : 0.f);
    float l = 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        mma_lohi(tS, desc_lo(smem_u32(sQ), 128) + k * 16, desc_hi(512), desc_lo(smem_u32(sK), 128) + k * 16, desc_hi(512),
                 idesc_s, k > 0);
      mma_commit(bars + 0);
    }
    for (int c = 0; c < nchunks; ++c) {
      // S of the NEXT chunk is issued before this chunk is processed (its TMEM buffer was drained one chunk ago)
      if (tid == 0 && c + 1 < nchunks) {
        const uint32_t tSn = tS + ((c + 1) & 1) * KC;
        const uint32_t kn_lo = desc_lo(smem_u32(sK) + (c + 1) * KC * (DH * 2), 128);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          mma_lohi(tSn, desc_lo(smem_u32(sQ), 128) + k * 16, desc_hi(512), kn_lo + k * 16, desc_hi(512), idesc_s, k > 0);
        mma_commit(bars + ((c + 1) & 1));
      }
      if (c & 1) { mbar_wait(bars + 1, ph_s1); ph_s1 ^= 1; } else { mbar_wait(bars + 0, ph_s0); ph_s0 ^= 1; }
      tc_fence_after();
      uint32_t s[32];
      tmem_ld_32x32(tS + (c & 1) * KC + lane_off + half * 32, s);
      tmem_wait_ld();
      const int k0 = c * KC + half * 32;  // first key of this thread's 32 columns
      uint32_t pk[16];
      if (has_bias) {
        const char* tabb = reinterpret_cast<const char*>(sTab + a_i);
        const int2* nbp = reinterpret_cast<const int2*>(sB + k0);
        if (k0 >= key_hi) {              // whole piece beyond the sequence (warp-uniform)
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        } else if (k0 + 32 <= key_hi) {  // fully valid piece: no per-key predicates
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const int2 nb = nbp[j >> 1];
            const float e0 = ex2_approx((__uint_as_float(s[j]) - m_i) + *reinterpret_cast<const float*>(tabb + nb.x));
            const float e1 = ex2_approx((__uint_as_float(s[j + 1]) - m_i) + *reinterpret_cast<const float*>(tabb + nb.y));
            l += e0 + e1;
            pk[j >> 1] = pack_bf16(e0, e1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const int2 nb = nbp[j >> 1];
            const float x0 = (__uint_as_float(s[j]) - m_i) + *reinterpret_cast<const float*>(tabb + nb.x);
            const float x1 = (__uint_as_float(s[j + 1]) - m_i) + *reinterpret_cast<const float*>(tabb + nb.y);
            const float e0 = (k0 + j < key_hi) ? ex2_approx(x0) : 0.f;
            const float e1 = (k0 + j + 1 < key_hi) ? ex2_approx(x1) : 0.f;
            l += e0 + e1;
            pk[j >> 1] = pack_bf16(e0, e1);
          }
        }
      } else if (__all_sync(0xffffffffu, k0 + 32 <= key_lo || k0 >= key_hi)) {
        // packed short sequences: this 32-key piece belongs to other sequences for every row of the warp
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = 0u;
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int ka = k0 + j;
          float e0 = (ka >= key_lo && ka < key_hi) ? ex2_approx(__uint_as_float(s[j]) - m_i) : 0.f;
          float e1 = (ka + 1 >= key_lo && ka + 1 < key_hi) ? ex2_approx(__uint_as_float(s[j + 1]) - m_i) : 0.f;
          l += e0 + e1;
          pk[j >> 1] = pack_bf16(e0, e1);
        }
      }
      if (c > 0) {  // the previous P V must have finished reading the P tile
        mbar_wait(bars + 2, ph_o);
        ph_o ^= 1;
      }
      // this thread's 32 columns of the P tile (128 x 64, K-major core matrices)
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8)
        *reinterpret_cast<uint4*>(sP + cm_off(rowt, half * 4 + c8, KC)) =
            make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        // O += P V_c     A: P [128 x 64] K-major ; B: V_c [N=32][K=64] MN-major (rows = keys)
        const uint32_t pv_lo = desc_lo(smem_u32(sP), 128), vv_lo = desc_lo(smem_u32(sV) + c * KC * (DH * 2), 512);
        const uint32_t acc_o = (c > 0) ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < KC / 16; ++k)
          mma_lohi(tO, pv_lo + k * 16, desc_hi((KC / 8) * 128), vv_lo + k * 64, desc_hi(128), idesc_o, k > 0 ? 1u : acc_o);
        mma_commit(bars + 2);
      }
    }
    // ---- epilogue of this query tile: each column half stores 16 of the 32 output columns
    sL[half * 128 + rowt] = l;
    mbar_wait(bars + 2, ph_o);
    ph_o ^= 1;
    tc_fence_after();
    __syncthreads();
    {
      const float lt = sL[rowt] + sL[128 + rowt];
      uint32_t o[16];
      tmem_ld_32x16(tO + lane_off + half * 16, o);
      tmem_wait_ld();
      if (valid) {
        const float inv_l = 1.f / lt;
        uint4* dst = reinterpret_cast<uint4*>(p.o + tok * p.ldo + head * DH + half * 16);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[8 * j + 0]) * inv_l, __uint_as_float(o[8 * j + 1]) * inv_l);
          u.y = pack_bf16(__uint_as_float(o[8 * j + 2]) * inv_l, __uint_as_float(o[8 * j + 3]) * inv_l);
          u.z = pack_bf16(__uint_as_float(o[8 * j + 4]) * inv_l, __uint_as_float(o[8 * j + 5]) * inv_l);
          u.w = pack_bf16(__uint_as_float(o[8 * j + 6]) * inv_l, __uint_as_float(o[8 * j + 7]) * inv_l);
          dst[j] = u;
        }
        if (half == 0 && p.lse != nullptr) p.lse[tok * p.heads + head] = m_i + log2f(lt);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  }  // blocks
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// =============================================================================================== backward
// One CTA per (sequence block, head), 128 threads. Q~ (= 8 log2e q^), dO, lse and delta = rowsum(dO*O) stay in
// shared memory for the whole block; key chunks of 128 rows stream through. Per (chunk c, query tile i):
//   S = Q~_i K^_c^T, dP = dO_i V_c^T               (TMEM, 128 columns each)
//   p = exp2(S + log2e*bias - lse), dz = p*(dP - delta)   -> smem as bf16 core matrices (P, dS)
//   dV_c += P^T dO_i ; dK^_c += dS^T Q~_i ; dQ_i += dS K^_c      (TMEM; dQ keeps one 32-column slot per query tile)
// One mbarrier: every tcgen05.commit covers all MMAs issued so far, so a single wait per step orders everything.
struct AttnBwdParams {
  AttnParams f;
  const __nv_bfloat16* d_o;
  __nv_bfloat16* dq;
  __nv_bfloat16* dkv;
  float* dq_scale;
  float* dk_scale;
  float* dbias_table;
};

constexpr int BKC = 128;  // keys per chunk (backward)
constexpr int BKH = 64;   // keys per sub-step: S / dP of the two 64-key halves have their own TMEM columns and barriers

// Pipeline of one (key chunk c, query tile i) step n — 256 threads, thread = (row t & 127, 32-column quarter of the
// 64-key half):
//   top      : Q~/dO rows of tile n+1 (prefetched to registers during step n-1) -> tile buffer (n+1) % 3
//   half A   : wait S_A/dP_A(n) -> tcgen05.ld -> barrier -> thread 0 issues S_A/dP_A(n+1) -> softmax / dS math
//   half B   : wait S_B/dP_B(n) -> tcgen05.ld -> math
//   tail     : wait dV/dK/dQ(n-1) retired -> P, dS tiles to shared memory -> barrier -> thread 0 issues
//              S_B/dP_B(n+1), then dV_c += P^T dO_i, dK^_c += dS^T Q~_i, dQ_i += dS K^_c
// so the tensor pipe works on S/dP of the next tile and on the three gradient MMAs of this tile while the CUDA cores
// do the element-wise work; TMEM: S 128 + dP 128 + dV 32 + dK 32 + dQ 32 x (tiles <= 6) = 512 columns.
__global__ void __launch_bounds__(256, 1)
attn_bwd_kernel(const AttnBwdParams bp) {
  const AttnParams& p = bp.f;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int head = blockIdx.x % p.heads;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowt = tid & 127, half = tid >> 7;
  const int R = p.ns * p.nst;
  const int nchunks = (R + BKC - 1) / BKC;
  const int ntiles = (R + QT - 1) / QT;
  const bool has_bias = p.bias_table != nullptr;
  const int tab_n = has_bias ? (2 * p.gh - 1) * (2 * p.gw - 1) : 0;   // words of the table in global memory
  const int tab_s = has_bias ? tab_words(p.gh, p.gw) : 0;              // words of a padded copy in shared memory
  const int tpitch = tab_pitch(p.gw), tw = 2 * p.gw - 1;
  const int r_lse = (R + 31) & ~31;                                    // rows that own an lse / delta slot

  constexpr uint32_t TILE_B = QT * DH * 2;         // bytes of one Q~ / dO tile buffer
  uint8_t* sQt = smem;                             // 3 x (128 x 32)  Q~ tile ring
  uint8_t* sdOt = sQt + 3 * TILE_B;                // 3 x (128 x 32)  dO tile ring
  uint8_t* sK = sdOt + 3 * TILE_B;                 // 128 x 32
  uint8_t* sV = sK + BKC * DH * 2;                 // 128 x 32
  uint8_t* sP = sV + BKC * DH * 2;                 // 128 x 128
  uint8_t* sdS = sP + QT * BKC * 2;                // 128 x 128
  short* sBn = reinterpret_cast<short*>(sdS + QT * BKC * 2);  // [r_pad] -4 * B_j: byte offset of key j inside the bias table
  float* sLse = reinterpret_cast<float*>(sBn + p.r_pad);       // (16-bit: the padded table is < 32 KB; r_pad is a multiple of 128)
  float* sDelta = sLse + r_lse;
  float* sScale = sDelta + r_lse;                  // [64]
  float* sRed = sScale + 64;                       // [64]
  float* sTab = sRed + 64;                         // [tab_s] bias * log2e, rows at the conflict-free pitch
  float* sdTab = sTab + tab_s;                     // [8][tab_s] per-warp private dbias accumulators
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sdTab + 8 * tab_s) + 15) & ~uintptr_t(15));
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  if (tid == 0) {
    mbar_init(bars + 0, 1);  // S_A / dP_A of the current step ready
    mbar_init(bars + 1, 1);  // S_B / dP_B of the current step ready
    mbar_init(bars + 2, 3);  // dV / dK / dQ MMAs of the current step retired (one commit per issuing thread): P, dS tiles free
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (tid < DH) {
    sScale[tid] = p.q_scale[tid];
    sScale[DH + tid] = p.k_scale[tid];
  }
  if (tid < 64) sRed[tid] = 0.f;
  for (int i = tid; i < tab_n; i += blockDim.x)
    sTab[(i / tw) * tpitch + i % tw] = p.bias_table[(long long)head * tab_n + i] * kLog2e;
  for (int i = tid; i < 8 * tab_s; i += blockDim.x) sdTab[i] = 0.f;
  for (int r = tid; r < p.r_pad; r += blockDim.x) {
    const int kp = r % p.n;
    sBn[r] = (short)(-4 * ((kp / p.gw) * tpitch + (kp % p.gw)));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 288, tdQ = tmem + 320;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  constexpr uint32_t idesc_s = make_idesc_bf16(QT, BKH, false, false);   // [128 q] x [64 keys], K = 32
  constexpr uint32_t idesc_kv = make_idesc_bf16(BKC, DH, true, true);    // [128 keys] x [32], K = 128 queries
  constexpr uint32_t idesc_q = make_idesc_bf16(QT, DH, false, true);     // [128 q] x [32], K = 128 keys
  constexpr uint32_t P_RS = (BKC / 8) * 128;                             // row-group stride of the P / dS tiles
  constexpr uint32_t KH_B = BKH * DH * 2;                                // bytes of 64 key rows of sK / sV
  float* my_dtab = sdTab + warp * tab_s;
  uint32_t ph_a = 0, ph_b = 0, ph_m = 0;
  bool mma_pending = false;
  float acc_qs[DH], acc_ks[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) acc_qs[d] = acc_ks[d] = 0.f;
  int gstep = 0;  // running step counter -> tile ring slot

  // persistent over the sequence blocks of this head: tables, barriers, TMEM and the dbias / scale-gradient
  // accumulators are set up (and flushed) once per CTA
  for (int blk = blockIdx.x / p.heads; blk < p.num_blocks; blk += gridDim.x / p.heads) {
  // ---- lse and delta = rowsum(dO * O) for the whole block
  for (int r = tid; r < r_lse; r += blockDim.x) {
    bool valid = false;
    const long long tok = (r < R) ? row_token(p, blk, r, valid) : 0;
    if (valid) {
      float go[DH], oo[DH];
      load_head_row(bp.d_o + tok * p.ldo + head * DH, go);
      load_head_row(p.o + tok * p.ldo + head * DH, oo);
      float dl = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) dl = fmaf(go[d], oo[d], dl);
      sDelta[r] = dl;
      sLse[r] = p.lse[tok * p.heads + head];
    } else {
      sDelta[r] = 0.f;
      sLse[r] = INFINITY;  // exp2(x - inf) = 0: invalid query rows contribute nothing
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // half 0 builds Q~ rows, half 1 copies dO rows. `pre` holds the raw global row of a tile not yet in shared memory.
  uint4 pre[4];
  auto prefetch_tile_row = [&](int tile) {
    const int r = tile * QT + rowt;
    bool valid = false;
    const long long tok = (r < R) ? row_token(p, blk, r, valid) : 0;
    if (valid) {
      const uint4* src = reinterpret_cast<const uint4*>((half == 0 ? p.q + tok * p.ldq : bp.d_o + tok * p.ldo) + head * DH);
#pragma unroll
      for (int j = 0; j < 4; ++j) pre[j] = src[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) pre[j] = make_uint4(0, 0, 0, 0);
    }
  };
  auto store_tile_row = [&](int buf) {
    if (half == 0) {
      float v[DH];
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w4[4] = {pre[j].x, pre[j].y, pre[j].z, pre[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[8 * j + 2 * k] = bf16_lo(w4[k]);
          v[8 * j + 2 * k + 1] = bf16_hi(w4[k]);
        }
      }
#pragma unroll
      for (int d = 0; d < DH; ++d) ss = fmaf(v[d], v[d], ss);
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int d = 0; d < DH; ++d) v[d] *= inv * sScale[d] * (kScale * kLog2e);
      store_row_cm(sQt + buf * TILE_B, rowt, v);
    } else {
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) *reinterpret_cast<uint4*>(sdOt + buf * TILE_B + cm_off(rowt, c8, DH)) = pre[c8];
    }
  };
  // S_h = Q~ K^_h^T and dP_h = dO V_h^T for the 64-key half h of the chunk (thread 0 only)
  auto issue_s = [&](int buf, int h) {
    const uint32_t q_lo = desc_lo(smem_u32(sQt) + buf * TILE_B, 128), o_lo = desc_lo(smem_u32(sdOt) + buf * TILE_B, 128);
    const uint32_t k_lo = desc_lo(smem_u32(sK) + h * KH_B, 128), v_lo = desc_lo(smem_u32(sV) + h * KH_B, 128);
#pragma unroll
    for (int k = 0; k < DH / 16; ++k) {
      mma_lohi(tS + h * BKH, q_lo + k * 16, desc_hi(512), k_lo + k * 16, desc_hi(512), idesc_s, k > 0);
      mma_lohi(tdP + h * BKH, o_lo + k * 16, desc_hi(512), v_lo + k * 16, desc_hi(512), idesc_s, k > 0);
    }
    mma_commit(bars + h);
  };

  for (int c = 0; c < nchunks; ++c) {
    // ---- K^_c (half 0) and V_c (half 1); thread = key row. Every MMA that reads sK / sV has retired (chunk epilogue).
    const int kr = c * BKC + rowt;
    bool kvalid = false;
    const long long ktok = (kr < R) ? row_token(p, blk, kr, kvalid) : 0;
    float kraw[DH];
    float kinv = 0.f;
    // Global latency off the chunk-start critical path: the first tile's rows were requested two steps ago (see the step
    // loop) or are requested here BEFORE the K / V rows are waited for, and the lines the next chunk start / the next
    // sequence block will need are pulled into L2 now (no registers held).
    if (c == 0 || ntiles < 2) prefetch_tile_row(0);
    {
      const int nblk = blk + gridDim.x / p.heads;
      const bool next_chunk = c + 1 < nchunks;
      if (next_chunk || nblk < p.num_blocks) {
        const int nr = next_chunk ? kr + BKC : rowt;
        bool nvalid = false;
        const long long ntok = (nr < R) ? row_token(p, next_chunk ? blk : nblk, nr, nvalid) : 0;
        if (nvalid) {
          const __nv_bfloat16* a = p.kv + ntok * p.ldkv + (half ? p.inner : 0) + head * DH;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
      if (!next_chunk && nblk < p.num_blocks) {   // last chunk: the next block's dO / O rows (delta prologue) and lse
        for (int r = tid; r < R; r += blockDim.x) {
          bool nvalid = false;
          const long long ntok = row_token(p, nblk, r, nvalid);
          if (nvalid) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(bp.d_o + ntok * p.ldo + head * DH));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.o + ntok * p.ldo + head * DH));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.q + ntok * p.ldq + head * DH));
          }
        }
      }
    }
    if (half == 0) {
      if (kvalid) {
        kinv = load_head_row(p.kv + ktok * p.ldkv + head * DH, kraw);
        float v[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) v[d] = kraw[d] * kinv * sScale[DH + d];
        store_row_cm(sK, rowt, v);
      } else {
        store_zero_row_cm(sK, rowt);
      }
    } else {
      if (kvalid) {
        const uint4* s4 = reinterpret_cast<const uint4*>(p.kv + ktok * p.ldkv + p.inner + head * DH);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) *reinterpret_cast<uint4*>(sV + cm_off(rowt, c8, DH)) = s4[c8];
      } else {
        store_zero_row_cm(sV, rowt);
      }
    }
    store_tile_row(gstep % 3);
    if (ntiles > 1) prefetch_tile_row(1);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_s(gstep % 3, 0);
      issue_s(gstep % 3, 1);
    }
    for (int i = 0; i < ntiles; ++i, ++gstep) {
      const int buf = gstep % 3, nbuf = (gstep + 1) % 3;
      const int r = i * QT + rowt;
      const int my_seq = r / p.nst;
      const int my_pos = r - my_seq * p.nst;
      const int key_lo = my_seq * p.nst, key_hi = min(R, key_lo + p.n);
      const int a_i = (my_pos / p.gw + p.gh - 1) * tpitch + (my_pos % p.gw) + p.gw - 1;
      const float lse_i = r < r_lse ? sLse[r] : INFINITY, delta_i = r < r_lse ? sDelta[r] : 0.f;   // rows >= R: p = exp2(-inf) = 0
      // next tile's rows -> ring slot nbuf (last read by the MMAs of step n-2, retired before step n-1 wrote P / dS)
      if (i + 1 < ntiles) {
        store_tile_row(nbuf);
        if (i + 2 < ntiles) prefetch_tile_row(i + 2);
        else if (c + 1 < nchunks) prefetch_tile_row(0);   // first tile of the next chunk: in registers across the chunk's last step
      }
      uint32_t pk[32], dk[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // This thread's 32 columns of the 64-key half are two 16-column sub-pieces. With the bias table they lie 32
        // keys apart: their table offsets differ by >= 55 entries, more than the spread of a warp's 32 query positions
        // (<= 54), so the two read-modify-write chains of the dbias accumulation never touch the same word and can be
        // interleaved (two chains in flight instead of one). Without bias the sub-pieces are simply adjacent.
        const int col0 = h * BKH + (has_bias ? half * 16 : half * 32);
        const int col1 = col0 + (has_bias ? 32 : 16);
        const int k0 = c * BKC + col0, k1 = c * BKC + col1;
        if (h == 0) { mbar_wait(bars + 0, ph_a); ph_a ^= 1; } else { mbar_wait(bars + 1, ph_b); ph_b ^= 1; }
        tc_fence_after();
        uint32_t s0[16], s1[16], dp0[16], dp1[16];
        tmem_ld_32x16(tS + lane_off + col0, s0);
        tmem_ld_32x16(tS + lane_off + col1, s1);
        tmem_ld_32x16(tdP + lane_off + col0, dp0);
        tmem_ld_32x16(tdP + lane_off + col1, dp1);
        tmem_wait_ld();
        if (h == 0) {
          // S_A / dP_A columns are in registers everywhere after this barrier: the tensor pipe may overwrite them
          fence_proxy_async_smem();
          tc_fence_before();
          __syncthreads();
          if (tid == 0 && i + 1 < ntiles) {
            tc_fence_after();
            issue_s(nbuf, 0);
          }
        }
        uint32_t* pk0 = pk + h * 16;      // packed P of sub-piece 0 (8 words), sub-piece 1 at +8
        uint32_t* dk0 = dk + h * 16;
        if (has_bias) {
          const char* tabb = reinterpret_cast<const char*>(sTab + a_i);
          // Per-warp private dbias table. The lanes of a warp are consecutive query positions (distinct A_i), so one
          // warp instruction never touches an address twice; consecutive instructions of the warp do (lane l at key j+1
          // hits what lane l-1 hit at key j), which is ordered by the in-order LSU pipe of a converged warp: the
          // accesses are volatile (no compiler reordering) and unconditional (invalid rows / keys add an exact 0).
          volatile char* dtb = reinterpret_cast<volatile char*>(my_dtab + a_i);
          const int* nb0 = reinterpret_cast<const int*>(sBn + k0);     // two 16-bit key offsets per word
          const int* nb1 = reinterpret_cast<const int*>(sBn + k1);
          if (k0 >= key_hi) {               // both sub-pieces beyond the sequence (warp-uniform): nothing to do
#pragma unroll
            for (int j = 0; j < 16; ++j) pk0[j] = dk0[j] = 0u;
          } else if (k1 + 16 <= key_hi) {   // both sub-pieces fully valid: no per-key predicates, two chains interleaved
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const int wa = nb0[j >> 1], wb = nb1[j >> 1];
              const int2 na = make_int2((int)(short)wa, wa >> 16), nb = make_int2((int)(short)wb, wb >> 16);
              const float pa0 = ex2_approx((__uint_as_float(s0[j]) - lse_i) + *reinterpret_cast<const float*>(tabb + na.x));
              const float pb0 = ex2_approx((__uint_as_float(s1[j]) - lse_i) + *reinterpret_cast<const float*>(tabb + nb.x));
              const float pa1 = ex2_approx((__uint_as_float(s0[j + 1]) - lse_i) + *reinterpret_cast<const float*>(tabb + na.y));
              const float pb1 = ex2_approx((__uint_as_float(s1[j + 1]) - lse_i) + *reinterpret_cast<const float*>(tabb + nb.y));
              const float da0 = pa0 * (__uint_as_float(dp0[j]) - delta_i);
              const float db0 = pb0 * (__uint_as_float(dp1[j]) - delta_i);
              const float da1 = pa1 * (__uint_as_float(dp0[j + 1]) - delta_i);
              const float db1 = pb1 * (__uint_as_float(dp1[j + 1]) - delta_i);
              volatile float* qa0 = reinterpret_cast<volatile float*>(dtb + na.x);
              volatile float* qb0 = reinterpret_cast<volatile float*>(dtb + nb.x);
              const float ta0 = *qa0, tb0 = *qb0;     // loads of both chains first, then both stores
              *qa0 = ta0 + da0;
              *qb0 = tb0 + db0;
              volatile float* qa1 = reinterpret_cast<volatile float*>(dtb + na.y);
              volatile float* qb1 = reinterpret_cast<volatile float*>(dtb + nb.y);
              const float ta1 = *qa1, tb1 = *qb1;
              *qa1 = ta1 + da1;
              *qb1 = tb1 + db1;
              pk0[j >> 1] = pack_bf16(pa0, pa1);
              pk0[8 + (j >> 1)] = pack_bf16(pb0, pb1);
              dk0[j >> 1] = pack_bf16(da0, da1);
              dk0[8 + (j >> 1)] = pack_bf16(db0, db1);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int kq = q ? k1 : k0;
              const int* nbp = q ? nb1 : nb0;
              const uint32_t* sq = q ? s1 : s0;
              const uint32_t* dq_ = q ? dp1 : dp0;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const int wn = nbp[j >> 1];
                const int2 nb = make_int2((int)(short)wn, wn >> 16);
                const float x0 = (__uint_as_float(sq[j]) - lse_i) + *reinterpret_cast<const float*>(tabb + nb.x);
                const float x1 = (__uint_as_float(sq[j + 1]) - lse_i) + *reinterpret_cast<const float*>(tabb + nb.y);
                const float p0 = (kq + j < key_hi) ? ex2_approx(x0) : 0.f;
                const float p1 = (kq + j + 1 < key_hi) ? ex2_approx(x1) : 0.f;
                const float d0 = p0 * (__uint_as_float(dq_[j]) - delta_i);
                const float d1 = p1 * (__uint_as_float(dq_[j + 1]) - delta_i);
                volatile float* q0 = reinterpret_cast<volatile float*>(dtb + nb.x);
                *q0 = *q0 + d0;
                volatile float* q1 = reinterpret_cast<volatile float*>(dtb + nb.y);
                *q1 = *q1 + d1;
                pk0[q * 8 + (j >> 1)] = pack_bf16(p0, p1);
                dk0[q * 8 + (j >> 1)] = pack_bf16(d0, d1);
              }
            }
          }
        } else if (__all_sync(0xffffffffu, k0 + 32 <= key_lo || k0 >= key_hi)) {
          // packed short sequences: this 32-key piece belongs to other sequences for every row of the warp
#pragma unroll
          for (int j = 0; j < 16; ++j) pk0[j] = dk0[j] = 0u;
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int kq = q ? k1 : k0;
            const uint32_t* sq = q ? s1 : s0;
            const uint32_t* dq_ = q ? dp1 : dp0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const int ka = kq + j;
              const float p0 = (ka >= key_lo && ka < key_hi) ? ex2_approx(__uint_as_float(sq[j]) - lse_i) : 0.f;
              const float p1 = (ka + 1 >= key_lo && ka + 1 < key_hi) ? ex2_approx(__uint_as_float(sq[j + 1]) - lse_i) : 0.f;
              const float d0 = p0 * (__uint_as_float(dq_[j]) - delta_i);
              const float d1 = p1 * (__uint_as_float(dq_[j + 1]) - delta_i);
              pk0[q * 8 + (j >> 1)] = pack_bf16(p0, p1);
              dk0[q * 8 + (j >> 1)] = pack_bf16(d0, d1);
            }
          }
        }
      }
      // the previous step's dV / dK / dQ MMAs must have finished reading the P / dS tiles
      if (mma_pending) {
        mbar_wait(bars + 2, ph_m);
        ph_m ^= 1;
      }
      // this thread's 4 x 16 columns of the P / dS tiles (K-major core matrices): half-chunk h, sub-piece q
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int colq = h * BKH + (has_bias ? q * 32 + half * 16 : half * 32 + q * 16);
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            const int cg = (colq >> 3) + c8;      // 8-column group inside the 128-key chunk
            const int w = h * 16 + q * 8 + 4 * c8;
            *reinterpret_cast<uint4*>(sP + cm_off(rowt, cg, BKC)) = make_uint4(pk[w], pk[w + 1], pk[w + 2], pk[w + 3]);
            *reinterpret_cast<uint4*>(sdS + cm_off(rowt, cg, BKC)) = make_uint4(dk[w], dk[w + 1], dk[w + 2], dk[w + 3]);
          }
        }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      // The MMAs of this step are issued by four threads of four different warps (= four schedulers): the issue code of
      // one thread (descriptor arithmetic + 28 tcgen05.mma) would otherwise make its warp the straggler of every barrier.
      // The three gradient products use separate accumulators, so their relative order does not matter; each issuing thread
      // commits its own MMAs to bars[2] (3 arrivals per phase).
      if (tid == 0) {
        tc_fence_after();
        if (i + 1 < ntiles) issue_s(nbuf, 1);   // S_B / dP_B of the next tile
      } else if (tid == 32) {                   // dV_c += P^T dO_i   (A: P^T MN-major, B: dO_i MN-major)
        tc_fence_after();
        const uint32_t ob_lo = desc_lo(smem_u32(sdOt) + buf * TILE_B, 512), pt_lo = desc_lo(smem_u32(sP), P_RS);
        const uint32_t acc_kv = (i > 0) ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)      // reduction over the 128 queries of this tile
          mma_lohi(tdV, pt_lo + k * (2 * P_RS / 16), desc_hi(128), ob_lo + k * 64, desc_hi(128), idesc_kv, k > 0 ? 1u : acc_kv);
        mma_commit(bars + 2);
      } else if (tid == 64) {                   // dK^_c += dS^T Q~_i
        tc_fence_after();
        const uint32_t qb_lo = desc_lo(smem_u32(sQt) + buf * TILE_B, 512), st_lo = desc_lo(smem_u32(sdS), P_RS);
        const uint32_t acc_kv = (i > 0) ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          mma_lohi(tdK, st_lo + k * (2 * P_RS / 16), desc_hi(128), qb_lo + k * 64, desc_hi(128), idesc_kv, k > 0 ? 1u : acc_kv);
        mma_commit(bars + 2);
      } else if (tid == 96) {                   // dQ_i += dS K^_c
        tc_fence_after();
        const uint32_t sq_lo = desc_lo(smem_u32(sdS), 128), kq_lo = desc_lo(smem_u32(sK), 512);
        const uint32_t acc_q = (c > 0) ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < BKC / 16; ++k)     // reduction over the 128 keys of this chunk
          mma_lohi(tdQ + i * DH, sq_lo + k * 16, desc_hi(P_RS), kq_lo + k * 64, desc_hi(128), idesc_q, k > 0 ? 1u : acc_q);
        mma_commit(bars + 2);
      }
      mma_pending = true;
    }
    // ---- chunk epilogue: wait for the last dV / dK MMAs; half 1 stores dV_c rows, half 0 stores dK_c rows
    mbar_wait(bars + 2, ph_m);
    ph_m ^= 1;
    mma_pending = false;
    tc_fence_after();
    {
      uint32_t a[32];
      tmem_ld_32x32((half ? tdV : tdK) + lane_off, a);
      tmem_wait_ld();
      if (kvalid && half == 1) {
        uint4* dv = reinterpret_cast<uint4*>(bp.dkv + ktok * p.ldkv + p.inner + head * DH);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dv[j] = make_uint4(pack_bf16(__uint_as_float(a[8 * j + 0]), __uint_as_float(a[8 * j + 1])),
                             pack_bf16(__uint_as_float(a[8 * j + 2]), __uint_as_float(a[8 * j + 3])),
                             pack_bf16(__uint_as_float(a[8 * j + 4]), __uint_as_float(a[8 * j + 5])),
                             pack_bf16(__uint_as_float(a[8 * j + 6]), __uint_as_float(a[8 * j + 7])));
      }
      if (kvalid && half == 0) {
        // k^ = ks * k / |k| : dk = (g - kbar (kbar.g)) / |k|, g = ks * dk^ ; dks += dk^ * kbar
        float g[DH], dot = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) {
          const float dkh = __uint_as_float(a[d]) * (1.f / kLog2e);
          const float kbar = kraw[d] * kinv;
          acc_ks[d] = fmaf(dkh, kbar, acc_ks[d]);
          g[d] = dkh * sScale[DH + d];
          dot = fmaf(g[d], kbar, dot);
        }
        uint4* dkp = reinterpret_cast<uint4*>(bp.dkv + ktok * p.ldkv + head * DH);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = (g[8 * j + e] - kraw[8 * j + e] * kinv * dot) * kinv;
          dkp[j] = make_uint4(pack_bf16(o8[0], o8[1]), pack_bf16(o8[2], o8[3]), pack_bf16(o8[4], o8[5]),
                              pack_bf16(o8[6], o8[7]));
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // all TMEM reads of dV / dK done before the next chunk's first MMA overwrites them
  }

  // ---- dQ epilogue (thread = query row; tile i handled by column-half i & 1)
  tc_fence_after();
  for (int i = half; i < ntiles; i += 2) {
    const int r = i * QT + rowt;
    bool valid = false;
    const long long tok = (r < R) ? row_token(p, blk, r, valid) : 0;
    uint32_t a[32];
    tmem_ld_32x32(tdQ + i * DH + lane_off, a);
    tmem_wait_ld();
    if (valid) {
      float qraw[DH];
      const float qinv = load_head_row(p.q + tok * p.ldq + head * DH, qraw);
      float g[DH], dot = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        const float dqh = __uint_as_float(a[d]) * kScale;  // d/dq^ = 8 * dz K^
        const float qbar = qraw[d] * qinv;
        acc_qs[d] = fmaf(dqh, qbar, acc_qs[d]);
        g[d] = dqh * sScale[d];
        dot = fmaf(g[d], qbar, dot);
      }
      uint4* dqp = reinterpret_cast<uint4*>(bp.dq + tok * p.ldq + head * DH);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = (g[8 * j + e] - qraw[8 * j + e] * qinv * dot) * qinv;
        dqp[j] = make_uint4(pack_bf16(o8[0], o8[1]), pack_bf16(o8[2], o8[3]), pack_bf16(o8[4], o8[5]),
                            pack_bf16(o8[6], o8[7]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  }  // blocks
  {
    // cross-thread reduction of the scale gradients through the (now idle) tile buffers: [256 threads][64 + 1] floats
    float* scratch = reinterpret_cast<float*>(smem);  // 66.5 KB <= tile rings + K + V + P (112 KB)
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      scratch[tid * 65 + d] = acc_qs[d];
      scratch[tid * 65 + DH + d] = acc_ks[d];
    }
    __syncthreads();
    if (tid < 64) {
      float t = 0.f;
      for (int r = 0; r < 256; ++r) t += scratch[r * 65 + tid];
      sRed[tid] = t;
    }
    __syncthreads();
  }
  if (tid < DH) {
    if (bp.dq_scale != nullptr) atomicAdd(bp.dq_scale + tid, sRed[tid]);
    if (bp.dk_scale != nullptr) atomicAdd(bp.dk_scale + tid, sRed[DH + tid]);
  }
  if (has_bias && bp.dbias_table != nullptr)
    for (int i = tid; i < tab_n; i += blockDim.x) {
      const int si = (i / tw) * tpitch + i % tw;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sdTab[w * tab_s + si];
      atomicAdd(bp.dbias_table + (long long)head * tab_n + i, t);
    }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// =============================================================================================== short sequences
// Sequences of n <= 32 tokens without an additive bias (the temporal attention of CTViT: n = t = 24, 4608 x 8
// (sequence, head) problems of 24 x 24 x 32 per volume batch). A 128-row tcgen05 tile would be > 80 % padding and the
// whole problem fits the registers of one warp, so here ONE WARP owns one (sequence, head): Q^, K^, V (dO, O) rows go
// straight from global memory into mma.sync m16n8k16 fragment images (bf16 operands, fp32 accumulators — the same operand
// rounding as the tcgen05 kernels), S / P / dS never leave registers and the transposed operands (V^T, K^T, P^T, dS^T,
// dO^T, Q^T) come from movmatrix. Same math as above: s = (8 log2e q^).k^, p = exp2(s - lse).
//
// Fragment images: thread (g = lane / 4, c = lane % 4) holds, for row block a (rows g + 8a), ONE 16-byte group = the eight
// head-dim columns 8c .. 8c+7 (a single coalesced load / store per row). The tensor-core k / n index of a head-dim column
// is therefore a fixed permutation of the real one — harmless, every operand uses the same permutation and the output
// fragments (columns 2c + 8 nt + e in MMA space) land on the thread's own eight real columns 8c + 2 nt + e.
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {  // 8x8 bf16 block image -> image of its transpose
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ uint32_t comp(const uint4& u, int b) { return b == 0 ? u.x : b == 1 ? u.y : b == 2 ? u.z : u.w; }
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
  v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
}
// l2-normalise one head row spread over the four threads of a quad, scale per column, pack to bf16; returns 1 / |row|
__device__ __forceinline__ float norm_row(const uint4& raw, const float (&sc)[8], uint4& img) {
  float v[8];
  unpack8(raw, v);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
  ss = quad_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  img.x = pack_bf16(v[0] * inv * sc[0], v[1] * inv * sc[1]);
  img.y = pack_bf16(v[2] * inv * sc[2], v[3] * inv * sc[3]);
  img.z = pack_bf16(v[4] * inv * sc[4], v[5] * inv * sc[5]);
  img.w = pack_bf16(v[6] * inv * sc[6], v[7] * inv * sc[7]);
  return inv;
}
// token of row r of sequence `seq`: spatial sequences are contiguous, temporal ones have stride h*w
__device__ __forceinline__ void seq_origin(const AttnParams& p, long long seq, long long& tok0, long long& tstride) {
  if (p.mode == 0) { tok0 = seq * p.n; tstride = 1; return; }
  const int hw = p.gh * p.gw;
  const long long b = seq / hw;
  tok0 = b * p.gt * hw + (seq - b * hw);
  tstride = hw;
}

// NB = number of 8-row blocks (ceil(n / 8)); rows / keys beyond n are zero images and masked probabilities
template <int NB>
__global__ void __launch_bounds__(256)
attn_short_fwd_kernel(const AttnParams p) {
  constexpr int MT = (NB + 1) / 2, NA = 2 * MT;
  __shared__ float sScale[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  if (tid < DH) {
    sScale[tid] = p.q_scale[tid] * (kScale * kLog2e);
    sScale[DH + tid] = p.k_scale[tid];
  }
  __syncthreads();
  float qs[8], ks[8], cmax = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { qs[i] = sScale[8 * c + i]; ks[i] = sScale[DH + 8 * c + i]; }
  for (int d = 0; d < DH; ++d) cmax = fmaxf(cmax, fabsf(sScale[d] * sScale[DH + d]));
  const long long pairs = (long long)p.num_seqs * p.heads;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (long long pair = (long long)blockIdx.x * 8 + warp; pair < pairs; pair += (long long)gridDim.x * 8) {
    const int head = (int)(pair % p.heads);
    long long tok0, tstride;
    seq_origin(p, pair / p.heads, tok0, tstride);
    uint4 qi[NA], ki[NA], vi[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int r = g + 8 * a;
      qi[a] = ki[a] = vi[a] = zero4;
      if (a < NB && r < p.n) {
        const long long tok = tok0 + r * tstride;
        qi[a] = __ldg(reinterpret_cast<const uint4*>(p.q + tok * p.ldq + head * DH + 8 * c));
        ki[a] = __ldg(reinterpret_cast<const uint4*>(p.kv + tok * p.ldkv + head * DH + 8 * c));
        vi[a] = __ldg(reinterpret_cast<const uint4*>(p.kv + tok * p.ldkv + p.inner + head * DH + 8 * c));
      }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      norm_row(qi[a], qs, qi[a]);
      norm_row(ki[a], ks, ki[a]);
    }
    // S = Q~ K^T
    float sacc[MT][NA][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NA; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[mt][nt][e] = 0.f;
        if (nt < NB) {
#pragma unroll
          for (int kd = 0; kd < 2; ++kd)
            mma_16816(sacc[mt][nt], comp(qi[2 * mt], 2 * kd), comp(qi[2 * mt + 1], 2 * kd), comp(qi[2 * mt], 2 * kd + 1),
                      comp(qi[2 * mt + 1], 2 * kd + 1), comp(ki[nt], 2 * kd), comp(ki[nt], 2 * kd + 1));
        }
      }
    // p = exp2(s - cmax) (never overflows: |s| <= cmax), masked beyond the sequence; row sums over the quad
    uint32_t pimg[NA][NA];
    float l[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) l[a] = 0.f;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NA; ++nt) {
        const int j0 = 8 * nt + 2 * c;
        const bool v0 = (nt < NB) && (j0 < p.n), v1 = (nt < NB) && (j0 + 1 < p.n);
        const float e0 = v0 ? ex2_approx(sacc[mt][nt][0] - cmax) : 0.f;
        const float e1 = v1 ? ex2_approx(sacc[mt][nt][1] - cmax) : 0.f;
        const float e2 = v0 ? ex2_approx(sacc[mt][nt][2] - cmax) : 0.f;
        const float e3 = v1 ? ex2_approx(sacc[mt][nt][3] - cmax) : 0.f;
        l[2 * mt] += e0 + e1;
        l[2 * mt + 1] += e2 + e3;
        pimg[2 * mt][nt] = pack_bf16(e0, e1);
        pimg[2 * mt + 1][nt] = pack_bf16(e2, e3);
      }
#pragma unroll
    for (int a = 0; a < NA; ++a) l[a] = quad_sum(l[a]);
    // O = P V  (B operand = V^T blocks)
    float oacc[MT][4][4];
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      uint32_t vt[NA];
#pragma unroll
      for (int jb = 0; jb < NA; ++jb) vt[jb] = movm_t(comp(vi[jb], nd));
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[mt][nd][e] = 0.f;
#pragma unroll
        for (int kj = 0; kj < MT; ++kj)
          mma_16816(oacc[mt][nd], pimg[2 * mt][2 * kj], pimg[2 * mt + 1][2 * kj], pimg[2 * mt][2 * kj + 1],
                    pimg[2 * mt + 1][2 * kj + 1], vt[2 * kj], vt[2 * kj + 1]);
      }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int r = g + 8 * a;
      if (a < NB && r < p.n) {
        const long long tok = tok0 + r * tstride;
        const float il = 1.f / l[a];
        const int mt = a >> 1, h2 = (a & 1) * 2;
        uint4 u;
        u.x = pack_bf16(oacc[mt][0][h2] * il, oacc[mt][0][h2 + 1] * il);
        u.y = pack_bf16(oacc[mt][1][h2] * il, oacc[mt][1][h2 + 1] * il);
        u.z = pack_bf16(oacc[mt][2][h2] * il, oacc[mt][2][h2 + 1] * il);
        u.w = pack_bf16(oacc[mt][3][h2] * il, oacc[mt][3][h2 + 1] * il);
        *reinterpret_cast<uint4*>(p.o + tok * p.ldo + head * DH + 8 * c) = u;
        if (c == 0 && p.lse != nullptr) p.lse[tok * p.heads + head] = cmax + log2f(l[a]);
      }
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(256, 1)
attn_short_bwd_kernel(const AttnBwdParams bp) {
  const AttnParams& p = bp.f;
  constexpr int MT = (NB + 1) / 2, NA = 2 * MT;
  __shared__ float sScale[64];
  __shared__ float sRed[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  if (tid < DH) {
    sScale[tid] = p.q_scale[tid];
    sScale[DH + tid] = p.k_scale[tid];
  }
  if (tid < 64) sRed[tid] = 0.f;
  __syncthreads();
  float qsr[8], ksr[8], qs[8];   // raw scales, and the folded 8 log2e q_scale of Q~
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    qsr[i] = sScale[8 * c + i];
    ksr[i] = sScale[DH + 8 * c + i];
    qs[i] = qsr[i] * (kScale * kLog2e);
  }
  float accq[8], acck[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) accq[i] = acck[i] = 0.f;
  const long long pairs = (long long)p.num_seqs * p.heads;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (long long pair = (long long)blockIdx.x * 8 + warp; pair < pairs; pair += (long long)gridDim.x * 8) {
    const int head = (int)(pair % p.heads);
    long long tok0, tstride;
    seq_origin(p, pair / p.heads, tok0, tstride);
    uint4 qr[NA], kr[NA], vi[NA], gi[NA], qi[NA], ki[NA];   // raw q / k rows, V, dO, Q~, K^ images
    float delta[NA], lse[NA], qinv[NA], kinv[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int r = g + 8 * a;
      qr[a] = kr[a] = vi[a] = gi[a] = zero4;
      delta[a] = 0.f;
      lse[a] = INFINITY;   // exp2(s - inf) = 0 for rows beyond the sequence
      if (a < NB && r < p.n) {
        const long long tok = tok0 + r * tstride;
        qr[a] = __ldg(reinterpret_cast<const uint4*>(p.q + tok * p.ldq + head * DH + 8 * c));
        kr[a] = __ldg(reinterpret_cast<const uint4*>(p.kv + tok * p.ldkv + head * DH + 8 * c));
        vi[a] = __ldg(reinterpret_cast<const uint4*>(p.kv + tok * p.ldkv + p.inner + head * DH + 8 * c));
        gi[a] = __ldg(reinterpret_cast<const uint4*>(bp.d_o + tok * p.ldo + head * DH + 8 * c));
        const uint4 ou = __ldg(reinterpret_cast<const uint4*>(p.o + tok * p.ldo + head * DH + 8 * c));
        lse[a] = p.lse[tok * p.heads + head];
        float go[8], oo[8];
        unpack8(gi[a], go);
        unpack8(ou, oo);
#pragma unroll
        for (int i = 0; i < 8; ++i) delta[a] = fmaf(go[i], oo[i], delta[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      delta[a] = quad_sum(delta[a]);
      qinv[a] = norm_row(qr[a], qs, qi[a]);
      kinv[a] = norm_row(kr[a], ksr, ki[a]);
    }
    // S = Q~ K^T and dP = dO V^T, then p = exp2(s - lse), dz = p (dP - delta) packed as A-operand images
    uint32_t pimg[NA][NA], dimg[NA][NA];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NA; ++nt) {
        float sa[4] = {0.f, 0.f, 0.f, 0.f}, da[4] = {0.f, 0.f, 0.f, 0.f};
        if (nt < NB) {
#pragma unroll
          for (int kd = 0; kd < 2; ++kd) {
            mma_16816(sa, comp(qi[2 * mt], 2 * kd), comp(qi[2 * mt + 1], 2 * kd), comp(qi[2 * mt], 2 * kd + 1),
                      comp(qi[2 * mt + 1], 2 * kd + 1), comp(ki[nt], 2 * kd), comp(ki[nt], 2 * kd + 1));
            mma_16816(da, comp(gi[2 * mt], 2 * kd), comp(gi[2 * mt + 1], 2 * kd), comp(gi[2 * mt], 2 * kd + 1),
                      comp(gi[2 * mt + 1], 2 * kd + 1), comp(vi[nt], 2 * kd), comp(vi[nt], 2 * kd + 1));
          }
        }
        const int j0 = 8 * nt + 2 * c;
        const bool v0 = (nt < NB) && (j0 < p.n), v1 = (nt < NB) && (j0 + 1 < p.n);
        const float p0 = v0 ? ex2_approx(sa[0] - lse[2 * mt]) : 0.f;
        const float p1 = v1 ? ex2_approx(sa[1] - lse[2 * mt]) : 0.f;
        const float p2 = v0 ? ex2_approx(sa[2] - lse[2 * mt + 1]) : 0.f;
        const float p3 = v1 ? ex2_approx(sa[3] - lse[2 * mt + 1]) : 0.f;
        pimg[2 * mt][nt] = pack_bf16(p0, p1);
        pimg[2 * mt + 1][nt] = pack_bf16(p2, p3);
        dimg[2 * mt][nt] = pack_bf16(p0 * (da[0] - delta[2 * mt]), p1 * (da[1] - delta[2 * mt]));
        dimg[2 * mt + 1][nt] = pack_bf16(p2 * (da[2] - delta[2 * mt + 1]), p3 * (da[3] - delta[2 * mt + 1]));
      }
    // ---- dQ^ = dS K^   (B operand = K^T blocks), then the l2norm / scale backward of q
    {
      float acc[MT][4][4];
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        uint32_t kt[NA];
#pragma unroll
        for (int jb = 0; jb < NA; ++jb) kt[jb] = movm_t(comp(ki[jb], nd));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nd][e] = 0.f;
#pragma unroll
          for (int kj = 0; kj < MT; ++kj)
            mma_16816(acc[mt][nd], dimg[2 * mt][2 * kj], dimg[2 * mt + 1][2 * kj], dimg[2 * mt][2 * kj + 1],
                      dimg[2 * mt + 1][2 * kj + 1], kt[2 * kj], kt[2 * kj + 1]);
        }
      }
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const int r = g + 8 * a;
        const int mt = a >> 1, h2 = (a & 1) * 2;
        float raw[8], gg[8], dot = 0.f;
        unpack8(qr[a], raw);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dqh = acc[mt][i >> 1][h2 + (i & 1)] * kScale;   // d/dq^ = 8 dz K^
          const float qbar = raw[i] * qinv[a];
          accq[i] = fmaf(dqh, qbar, accq[i]);
          gg[i] = dqh * qsr[i];
          dot = fmaf(gg[i], qbar, dot);
        }
        dot = quad_sum(dot);
        if (a < NB && r < p.n) {
          const long long tok = tok0 + r * tstride;
          float o8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = (gg[i] - raw[i] * qinv[a] * dot) * qinv[a];
          *reinterpret_cast<uint4*>(bp.dq + tok * p.ldq + head * DH + 8 * c) =
              make_uint4(pack_bf16(o8[0], o8[1]), pack_bf16(o8[2], o8[3]), pack_bf16(o8[4], o8[5]), pack_bf16(o8[6], o8[7]));
        }
      }
    }
    // ---- dV = P^T dO and dK^ = dS^T Q~   (A operands = transposed P / dS blocks, B operands = dO^T / Q~^T blocks)
    {
      uint32_t pt[NA][NA], dt[NA][NA];   // [key block][query block]
#pragma unroll
      for (int jb = 0; jb < NA; ++jb)
#pragma unroll
        for (int ib = 0; ib < NA; ++ib) {
          pt[jb][ib] = movm_t(pimg[ib][jb]);
          dt[jb][ib] = movm_t(dimg[ib][jb]);
        }
      float av[MT][4][4], ak[MT][4][4];
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        uint32_t gt[NA], qt[NA];
#pragma unroll
        for (int ib = 0; ib < NA; ++ib) {
          gt[ib] = movm_t(comp(gi[ib], nd));
          qt[ib] = movm_t(comp(qi[ib], nd));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) av[mt][nd][e] = ak[mt][nd][e] = 0.f;
#pragma unroll
          for (int kq = 0; kq < MT; ++kq) {
            mma_16816(av[mt][nd], pt[2 * mt][2 * kq], pt[2 * mt + 1][2 * kq], pt[2 * mt][2 * kq + 1],
                      pt[2 * mt + 1][2 * kq + 1], gt[2 * kq], gt[2 * kq + 1]);
            mma_16816(ak[mt][nd], dt[2 * mt][2 * kq], dt[2 * mt + 1][2 * kq], dt[2 * mt][2 * kq + 1],
                      dt[2 * mt + 1][2 * kq + 1], qt[2 * kq], qt[2 * kq + 1]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const int r = g + 8 * a;
        const int mt = a >> 1, h2 = (a & 1) * 2;
        // k^ = ks * k / |k| : dk = (g - kbar (kbar.g)) / |k|, g = ks * dk^ ; dks += dk^ * kbar
        float raw[8], gg[8], dot = 0.f;
        unpack8(kr[a], raw);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dkh = ak[mt][i >> 1][h2 + (i & 1)] * (1.f / kLog2e);
          const float kbar = raw[i] * kinv[a];
          acck[i] = fmaf(dkh, kbar, acck[i]);
          gg[i] = dkh * ksr[i];
          dot = fmaf(gg[i], kbar, dot);
        }
        dot = quad_sum(dot);
        if (a < NB && r < p.n) {
          const long long tok = tok0 + r * tstride;
          float o8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = (gg[i] - raw[i] * kinv[a] * dot) * kinv[a];
          *reinterpret_cast<uint4*>(bp.dkv + tok * p.ldkv + head * DH + 8 * c) =
              make_uint4(pack_bf16(o8[0], o8[1]), pack_bf16(o8[2], o8[3]), pack_bf16(o8[4], o8[5]), pack_bf16(o8[6], o8[7]));
          *reinterpret_cast<uint4*>(bp.dkv + tok * p.ldkv + p.inner + head * DH + 8 * c) =
              make_uint4(pack_bf16(av[mt][0][h2], av[mt][0][h2 + 1]), pack_bf16(av[mt][1][h2], av[mt][1][h2 + 1]),
                         pack_bf16(av[mt][2][h2], av[mt][2][h2 + 1]), pack_bf16(av[mt][3][h2], av[mt][3][h2 + 1]));
        }
      }
    }
  }
  // scale gradients: sum over the eight row groups of the warp, then over the warps of the CTA, one atomic per column
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int m = 4; m < 32; m <<= 1) {
      accq[i] += __shfl_xor_sync(0xffffffffu, accq[i], m);
      acck[i] += __shfl_xor_sync(0xffffffffu, acck[i], m);
    }
    if (g == 0) {
      atomicAdd(&sRed[8 * c + i], accq[i]);
      atomicAdd(&sRed[DH + 8 * c + i], acck[i]);
    }
  }
  __syncthreads();
  if (tid < DH) {
    if (bp.dq_scale != nullptr) atomicAdd(bp.dq_scale + tid, sRed[tid]);
    if (bp.dk_scale != nullptr) atomicAdd(bp.dk_scale + tid, sRed[DH + tid]);
  }
}

// the short-sequence kernels apply to bias-free attention with n <= 32 (CTCLIP_ATTN_SHORT=0 forces the tcgen05 kernels)
bool use_short(const AttnParams& p) {
  const char* e = getenv("CTCLIP_ATTN_SHORT");   // read per call: the tests compare both paths in one process
  return !(e != nullptr && e[0] == '0') && p.bias_table == nullptr && p.n <= 32;
}
unsigned short_grid(const AttnParams& p, int ctas_per_sm) {
  const long long pairs = (long long)p.num_seqs * p.heads;
  long long ctas = (pairs + 7) / 8;
  const long long cap = (long long)ctclip::sm_count() * ctas_per_sm;
  return (unsigned)(ctas < cap ? ctas : cap);
}

size_t bwd_smem_bytes(const AttnParams& p) {
  const int tab_s = (p.bias_table != nullptr) ? tab_words(p.gh, p.gw) : 0;
  const int r_lse = (p.ns * p.nst + 31) & ~31;
  return (size_t)QT * DH * 2 * 6 + (size_t)BKC * DH * 2 * 2 + (size_t)QT * BKC * 2 * 2 + (size_t)p.r_pad * 2 +
         (size_t)r_lse * 4 * 2 + (size_t)tab_s * 4 * 9 + 128 * 4 + 64 + 16;
}

size_t fwd_smem_bytes(const AttnParams& p) {
  const int tab_n = (p.bias_table != nullptr) ? tab_words(p.gh, p.gw) + p.n : 0;
  const size_t kv_rows = (size_t)((p.ns * p.nst + KC - 1) / KC) * KC;
  return kv_rows * DH * 2 * 2 + QT * DH * 2 + QT * KC * 2 + (size_t)tab_n * 4 + 64 * 4 + 256 * 4 +
         (size_t)p.r_pad * 4 + 64 + 16;
}

int fill_params(AttnParams& p, const ctclip_attn_desc* d, const char* what) {
  if (d == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "%s: null descriptor", what);
  if (d->dim_head != DH) return ctclip::fail(CTCLIP_E_SHAPE, "%s: dim_head must be 32 (got %d)", what, d->dim_head);
  if (d->heads <= 0 || d->batch <= 0 || d->t <= 0 || d->h <= 0 || d->w <= 0)
    return ctclip::fail(CTCLIP_E_SHAPE, "%s: empty problem", what);
  p.heads = d->heads;
  p.inner = d->heads * DH;
  p.ldq = d->ldq; p.ldkv = d->ldkv; p.ldo = d->ldo;
  if ((p.ldq % 8) || (p.ldkv % 8) || (p.ldo % 8)) return ctclip::fail(CTCLIP_E_ALIGN, "%s: row strides must be multiples of 8", what);
  p.mode = d->temporal ? 1 : 0;
  p.gh = d->h; p.gw = d->w; p.gt = d->t;
  if (p.mode == 0) {
    p.n = d->h * d->w; p.nst = p.n; p.ns = 1; p.num_seqs = d->batch * d->t;
  } else {
    p.n = d->t;
    if (p.n > QT) return ctclip::fail(CTCLIP_E_SHAPE, "%s: temporal sequences longer than 128 are not supported", what);
    // short sequences are packed at a 32- / 64-row pitch: a warp (32 rows) then belongs to ONE sequence and touches only
    // the 32-key pieces of that sequence (everything else is skipped warp-uniformly)
    p.nst = p.n <= 32 ? 32 : (p.n <= 64 ? 64 : p.n);
    p.ns = QT / p.nst; p.num_seqs = d->batch * d->h * d->w;
  }
  p.r_pad = (p.ns * p.nst + 127) / 128 * 128;
  p.q_scale = d->q_scale; p.k_scale = d->k_scale;
  p.bias_table = d->temporal ? nullptr : d->bias_table;
  p.bias_rowmax = d->temporal ? nullptr : d->bias_rowmax;
  if (p.bias_table != nullptr && p.bias_rowmax == nullptr)
    return ctclip::fail(CTCLIP_E_SHAPE, "%s: bias_rowmax is required with bias_table", what);
  return ctclip::require_sm100();
}

}  // namespace

extern "C" int ctclip_attn_fwd(const ctclip_attn_desc* d, void* stream) {
  AttnParams p{};
  int rc = fill_params(p, d, "attn_fwd");
  if (rc) return rc;
  p.q = (const __nv_bfloat16*)d->q; p.kv = (const __nv_bfloat16*)d->kv; p.o = (__nv_bfloat16*)d->o; p.lse = d->lse;
  if (!p.q || !p.kv || !p.o) return ctclip::fail(CTCLIP_E_SHAPE, "attn_fwd: null pointer");
  if (use_short(p)) {
    const unsigned grid = short_grid(p, 2);
    switch ((p.n + 7) / 8) {
      case 1: attn_short_fwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(p); break;
      case 2: attn_short_fwd_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(p); break;
      case 3: attn_short_fwd_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(p); break;
      default: attn_short_fwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(p); break;
    }
    return ctclip::check_launch("attn_fwd(short)");
  }
  const size_t smem = fwd_smem_bytes(p);
  if (smem > 227 * 1024) return ctclip::fail(CTCLIP_E_SHAPE, "attn_fwd: sequence too long for shared memory (%zu B)", smem);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const long long blocks = ((long long)p.num_seqs + p.ns - 1) / p.ns;
  p.num_blocks = (int)blocks;
  // persistent: two CTAs per SM (shared memory / TMEM allow two), the same number of CTAs for every head
  long long per_head = (2LL * ctclip::sm_count()) / p.heads;
  if (per_head < 1) per_head = 1;
  if (per_head > blocks) per_head = blocks;
  attn_fwd_kernel<<<(unsigned)(per_head * p.heads), 256, smem, (cudaStream_t)stream>>>(p);
  return ctclip::check_launch("attn_fwd");
}

extern "C" int ctclip_attn_bwd(const ctclip_attn_desc* d, void* stream) {
  AttnBwdParams bp{};
  AttnParams& p = bp.f;
  int rc = fill_params(p, d, "attn_bwd");
  if (rc) return rc;
  p.q = (const __nv_bfloat16*)d->q; p.kv = (const __nv_bfloat16*)d->kv; p.o = (__nv_bfloat16*)d->o; p.lse = d->lse;
  bp.d_o = (const __nv_bfloat16*)d->d_o; bp.dq = (__nv_bfloat16*)d->dq; bp.dkv = (__nv_bfloat16*)d->dkv;
  bp.dq_scale = d->dq_scale; bp.dk_scale = d->dk_scale; bp.dbias_table = d->dbias_table;
  if (!p.q || !p.kv || !p.o || !p.lse || !bp.d_o || !bp.dq || !bp.dkv)
    return ctclip::fail(CTCLIP_E_SHAPE, "attn_bwd: null pointer");
  if (use_short(p)) {
    const unsigned grid = short_grid(p, 1);
    switch ((p.n + 7) / 8) {
      case 1: attn_short_bwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(bp); break;
      case 2: attn_short_bwd_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(bp); break;
      case 3: attn_short_bwd_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(bp); break;
      default: attn_short_bwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(bp); break;
    }
    return ctclip::check_launch("attn_bwd(short)");
  }
  if (p.r_pad / QT > 6) return ctclip::fail(CTCLIP_E_SHAPE, "attn_bwd: sequences longer than 768 tokens are not supported");
  const size_t smem = bwd_smem_bytes(p);
  if (smem > 227 * 1024) return ctclip::fail(CTCLIP_E_SHAPE, "attn_bwd: sequence too long for shared memory (%zu B)", smem);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const long long blocks = ((long long)p.num_seqs + p.ns - 1) / p.ns;
  p.num_blocks = (int)blocks;
  long long per_head = (long long)ctclip::sm_count() / p.heads;   // one CTA per SM (512 TMEM columns each)
  if (per_head < 1) per_head = 1;
  if (per_head > blocks) per_head = blocks;
  attn_bwd_kernel<<<(unsigned)(per_head * p.heads), 256, smem, (cudaStream_t)stream>>>(bp);
  return ctclip::check_launch("attn_bwd");
}
