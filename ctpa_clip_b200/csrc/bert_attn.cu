// Fused BertSelfAttention core (HF modeling_bert.py BertSelfAttention.forward: scores = Q K^T / sqrt(d) + key mask ->
// softmax -> attention-probs dropout -> P V) and its backward, for head dim 64, on packed [tokens, 3 * hidden] bf16 QKV.
//
// The text tower is 96 (batch, head) problems of 512 x 512 x 64 per layer: far below one tcgen05 wave and dominated by
// the fp32 score / probability round trips through HBM when done as separate GEMM + softmax launches (300 MB per layer
// each way). Here scores never leave registers (FlashAttention-2 tiling): warp-level mma.sync m16n8k16 (bf16 operands,
// fp32 accumulators), 64-row tiles staged in XOR-swizzled shared memory with cp.async, fragments via ldmatrix(.trans).
//   forward : CTA = 64 queries of one (b, h); streams 64-key blocks; online softmax in the exp2 domain; saves lse
//   backward: delta = rowsum(dO * O); dK/dV kernel: CTA = 64 keys, streams query blocks (S^T = K Q^T so the keys are the
//             accumulator rows); dQ kernel: CTA = 64 queries, streams key blocks. 7 small products instead of 5 — the
//             problem is bandwidth/latency-bound, not FLOP-bound, and nothing is reduced across CTAs (no atomics).
// Key blocks whose keys are all masked (padding) are skipped. Dropout is the stateless hash of bert.cu (same element
// index ((b H + h) L + i) L + j and seed), so the masks of forward and backward agree by construction.
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {
using namespace ptx;

constexpr int HD = 64;          // head dim
constexpr int BT = 64;          // tile rows (queries or keys)
constexpr int TILE_BYTES = BT * HD * 2;
constexpr float kLog2e = 1.4426950408889634f;

struct BertAttnParams {
  const __nv_bfloat16* qkv;     // [B*L][3*D]: q | k | v, head h at columns h*64
  const long long* mask;        // [B][L], 1 = attend
  __nv_bfloat16* out;           // [B*L][D]   context
  float* lse;                   // [B*H][L]   log2-domain log-sum-exp of the scaled, masked scores (+inf: no valid key)
  const __nv_bfloat16* dout;    // [B*L][D]
  __nv_bfloat16* dqkv;          // [B*L][3*D]
  float* delta;                 // [B*H][L]
  int B, H, L, D;
  float scale2;                 // log2(e) / sqrt(d)
  float scale;                  // 1 / sqrt(d)
  float p_drop;
  unsigned seed;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tile element (row, 16-byte chunk) -> byte offset; chunks are XOR-swizzled with the row so that ldmatrix (8 rows, same
// chunk) and the row-wise cp.async stores are both bank-conflict free
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// 64 x 64 bf16 tile: rows row0 .. row0+63 of one sequence (rows >= L are zero-filled), 64 columns at `col`; 128 threads
__device__ __forceinline__ void load_tile(uint32_t tile, const __nv_bfloat16* base, long long pitch, int row0, int L, int col) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + 128 * i;
    const int r = idx >> 3, ch = idx & 7;
    const bool ok = row0 + r < L;
    cp_async16(tile + tile_off(r, ch), base + (long long)(ok ? row0 + r : 0) * pitch + col + ch * 8, ok);
  }
}

// A fragments (16 rows x 64 k) of tile rows r0 .. r0+15: a[ks][0..3]
__device__ __forceinline__ void load_a_frags(uint32_t tile, int r0, uint32_t (&a)[4][4]) {
  const int lane = threadIdx.x & 31;
  const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8, chs = lane >> 4;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(tile + tile_off(row, 2 * ks + chs), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// C[16 x 64] (+)= A[16 x 64] * T^T, T = tile [64 n-rows][64 k]   (B operand "col-major" = the tile's rows as stored)
__device__ __forceinline__ void mma_a_tileT(float (&c)[8][4], const uint32_t (&a)[4][4], uint32_t tile) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int row = 8 * nt + (lane & 7);
    uint32_t b[8];
    ldsm_x4(tile + tile_off(row, lane >> 3), b[0], b[1], b[2], b[3]);
    ldsm_x4(tile + tile_off(row, 4 + (lane >> 3)), b[4], b[5], b[6], b[7]);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma16816(c[nt], a[ks], b[2 * ks], b[2 * ks + 1]);
  }
}

// C[16 x 64] += A[16 x 64(k = tile rows)] * T, T = tile [64 k-rows][64 n]   (B operand through ldmatrix.trans)
__device__ __forceinline__ void mma_a_tile(float (&c)[8][4], const uint32_t (&a)[4][4], uint32_t tile) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int row = 16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int np = 0; np < 4; ++np) {   // two n-tiles per ldmatrix
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile + tile_off(row, 2 * np + (lane >> 4)), b0, b1, b2, b3);
      mma16816(c[2 * np], a[ks], b0, b1);
      mma16816(c[2 * np + 1], a[ks], b2, b3);
    }
  }
}

// accumulator fragments [16 x 64] -> A fragments of the same matrix (bf16)
__device__ __forceinline__ void c_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_bf16(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf16(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf16(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf16(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

// [16 x 64] accumulator rows (g, g + 8 of the warp's 16 rows) -> bf16, 4-byte stores; thread columns 2c + 8 nt
__device__ __forceinline__ void store_rows(__nv_bfloat16* base, long long pitch, int row0, int L, int col, const float (&c)[8][4],
                                           float s0, float s1) {
  const int lane = threadIdx.x & 31, g = lane >> 2, cc = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (row0 + g < L)
      *reinterpret_cast<uint32_t*>(base + (long long)(row0 + g) * pitch + col + 8 * nt + 2 * cc) = pack_bf16(c[nt][0] * s0, c[nt][1] * s0);
    if (row0 + g + 8 < L)
      *reinterpret_cast<uint32_t*>(base + (long long)(row0 + g + 8) * pitch + col + 8 * nt + 2 * cc) =
          pack_bf16(c[nt][2] * s1, c[nt][3] * s1);
  }
}

// per-key additive mask (0 / -inf) of the sequence in shared memory + per-64-key-block "any valid key" flags
__device__ __forceinline__ void load_key_mask(const BertAttnParams& p, int b, float* s_mask, int* s_any, int nkb) {
  for (int j = threadIdx.x; j < nkb * BT; j += blockDim.x)
    s_mask[j] = (j < p.L && p.mask[(long long)b * p.L + j] != 0) ? 0.f : -INFINITY;
  __syncthreads();
  if (threadIdx.x < nkb) {
    int any = 0;
    for (int j = 0; j < BT; ++j) any |= (s_mask[threadIdx.x * BT + j] == 0.f);
    s_any[threadIdx.x] = any;
  }
  __syncthreads();
}

// =============================================================================================== forward
__global__ void __launch_bounds__(128, 4)
bert_attn_fwd_kernel(const BertAttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nkb = (p.L + BT - 1) / BT;
  uint8_t* tQ = smem;                     // 1 tile
  uint8_t* tK = tQ + TILE_BYTES;          // 2 tiles (double buffer)
  uint8_t* tV = tK + 2 * TILE_BYTES;      // 2 tiles
  float* s_mask = reinterpret_cast<float*>(tV + 2 * TILE_BYTES);
  int* s_any = reinterpret_cast<int*>(s_mask + nkb * BT);
  const int z = blockIdx.y, b = z / p.H, h = z - b * p.H;
  const int q0 = blockIdx.x * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, cc = lane & 3;
  const long long pitch = 3LL * p.D;
  const __nv_bfloat16* seq = p.qkv + (long long)b * p.L * pitch;
  load_tile(smem_u32(tQ), seq, pitch, q0, p.L, h * HD);
  cp_async_commit();
  load_key_mask(p, b, s_mask, s_any, nkb);
  // first valid key block
  int kb = 0;
  while (kb < nkb && !s_any[kb]) ++kb;
  if (kb < nkb) {
    load_tile(smem_u32(tK), seq, pitch, kb * BT, p.L, p.D + h * HD);
    load_tile(smem_u32(tV), seq, pitch, kb * BT, p.L, 2 * p.D + h * HD);
  }
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qa[4][4];
  load_a_frags(smem_u32(tQ), warp * 16, qa);
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int qi0 = q0 + warp * 16 + g, qi1 = qi0 + 8;
  const float keep_scale = 1.f / (1.f - p.p_drop);
  int buf = 0;
  while (kb < nkb) {
    int nxt = kb + 1;
    while (nxt < nkb && !s_any[nxt]) ++nxt;
    if (nxt < nkb) {   // prefetch the next valid key block into the other buffer
      load_tile(smem_u32(tK) + (buf ^ 1) * TILE_BYTES, seq, pitch, nxt * BT, p.L, p.D + h * HD);
      load_tile(smem_u32(tV) + (buf ^ 1) * TILE_BYTES, seq, pitch, nxt * BT, p.L, 2 * p.D + h * HD);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    mma_a_tileT(s, qa, smem_u32(tK) + buf * TILE_BYTES);
    // scaled, masked scores in the exp2 domain; running maxima of rows g / g + 8
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 mk = *reinterpret_cast<const float2*>(s_mask + kb * BT + 8 * nt + 2 * cc);
      s[nt][0] = fmaf(s[nt][0], p.scale2, mk.x); s[nt][1] = fmaf(s[nt][1], p.scale2, mk.y);
      s[nt][2] = fmaf(s[nt][2], p.scale2, mk.x); s[nt][3] = fmaf(s[nt][3], p.scale2, mk.y);
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float b0 = (mx0 == -INFINITY) ? 0.f : mx0, b1 = (mx1 == -INFINITY) ? 0.f : mx1;   // no valid key yet: exp2(-inf - 0) = 0
    const float r0 = ex2f(m0 - b0), r1 = ex2f(m1 - b1);
    m0 = mx0; m1 = mx1;
    l0 *= r0; l1 *= r1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= r0; o[nt][1] *= r0; o[nt][2] *= r1; o[nt][3] *= r1; }
    const unsigned long long e0 = ((unsigned long long)z * p.L + qi0) * p.L + kb * BT;
    const unsigned long long e1 = ((unsigned long long)z * p.L + qi1) * p.L + kb * BT;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float p0 = ex2f(s[nt][0] - b0), p1 = ex2f(s[nt][1] - b0), p2 = ex2f(s[nt][2] - b1), p3 = ex2f(s[nt][3] - b1);
      l0 += p0 + p1;
      l1 += p2 + p3;
      if (p.p_drop > 0.f) {
        const int j = 8 * nt + 2 * cc;
        p0 = keep_elem(e0 + j, p.seed, p.p_drop) ? p0 * keep_scale : 0.f;
        p1 = keep_elem(e0 + j + 1, p.seed, p.p_drop) ? p1 * keep_scale : 0.f;
        p2 = keep_elem(e1 + j, p.seed, p.p_drop) ? p2 * keep_scale : 0.f;
        p3 = keep_elem(e1 + j + 1, p.seed, p.p_drop) ? p3 * keep_scale : 0.f;
      }
      s[nt][0] = p0; s[nt][1] = p1; s[nt][2] = p2; s[nt][3] = p3;
    }
    uint32_t pa[4][4];
    c_to_a(s, pa);
    mma_a_tile(o, pa, smem_u32(tV) + buf * TILE_BYTES);
    __syncthreads();   // everyone is done with this buffer before the next prefetch overwrites it
    kb = nxt;
    buf ^= 1;
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  store_rows(p.out + (long long)b * p.L * p.D, p.D, q0 + warp * 16, p.L, h * HD, o, i0, i1);
  if (cc == 0 && p.lse != nullptr) {
    if (qi0 < p.L) p.lse[(long long)z * p.L + qi0] = l0 > 0.f ? m0 + log2f(l0) : INFINITY;
    if (qi1 < p.L) p.lse[(long long)z * p.L + qi1] = l1 > 0.f ? m1 + log2f(l1) : INFINITY;
  }
}

// =============================================================================================== backward
// delta[z][i] = sum_d dO[i][d] * O[i][d]; one warp per (token, head)
__global__ void __launch_bounds__(256)
bert_attn_delta_kernel(const BertAttnParams p) {
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long total = (long long)p.B * p.L * p.H;
  if (w >= total) return;
  const int lane = threadIdx.x & 31;
  const long long tok = w / p.H;
  const int h = (int)(w - tok * p.H);
  const uint32_t a = *reinterpret_cast<const uint32_t*>(p.dout + tok * p.D + h * HD + 2 * lane);
  const uint32_t o = *reinterpret_cast<const uint32_t*>(p.out + tok * p.D + h * HD + 2 * lane);
  float d = bf16_lo(a) * bf16_lo(o) + bf16_hi(a) * bf16_hi(o);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) d += __shfl_xor_sync(0xffffffffu, d, m);
  if (lane == 0) {
    const long long b = tok / p.L;
    p.delta[((long long)b * p.H + h) * p.L + (tok - b * p.L)] = d;
  }
}

// dK / dV of one 64-key block: rows of every accumulator are keys (S^T = K Q^T), query blocks stream through
__global__ void __launch_bounds__(128)
bert_attn_bwd_dkv_kernel(const BertAttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nqb = (p.L + BT - 1) / BT;
  uint8_t* tK = smem;
  uint8_t* tV = tK + TILE_BYTES;
  uint8_t* tQ = tV + TILE_BYTES;          // 2 tiles
  uint8_t* tG = tQ + 2 * TILE_BYTES;      // 2 tiles (dO)
  float* s_lse = reinterpret_cast<float*>(tG + 2 * TILE_BYTES);   // [nqb * 64]
  float* s_delta = s_lse + nqb * BT;
  const int z = blockIdx.y, b = z / p.H, h = z - b * p.H;
  const int k0 = blockIdx.x * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, cc = lane & 3;
  const long long pitch = 3LL * p.D;
  const __nv_bfloat16* seq = p.qkv + (long long)b * p.L * pitch;
  const __nv_bfloat16* gseq = p.dout + (long long)b * p.L * p.D;
  __nv_bfloat16* dseq = p.dqkv + (long long)b * p.L * pitch;
  const int kj0 = k0 + warp * 16 + g, kj1 = kj0 + 8;
  const bool kv0 = kj0 < p.L && p.mask[(long long)b * p.L + kj0] != 0;
  const bool kv1 = kj1 < p.L && p.mask[(long long)b * p.L + kj1] != 0;
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
  // a block of padding keys receives no probability mass: its gradients are zero
  int any = 0;
  for (int j = lane; j < BT; j += 32) any |= (k0 + j < p.L && p.mask[(long long)b * p.L + k0 + j] != 0);
  any = __syncthreads_or(any);
  if (any) {
    load_tile(smem_u32(tK), seq, pitch, k0, p.L, p.D + h * HD);
    load_tile(smem_u32(tV), seq, pitch, k0, p.L, 2 * p.D + h * HD);
    load_tile(smem_u32(tQ), seq, pitch, 0, p.L, h * HD);
    load_tile(smem_u32(tG), gseq, p.D, 0, p.L, h * HD);
    cp_async_commit();
    for (int i = threadIdx.x; i < nqb * BT; i += blockDim.x) {
      s_lse[i] = i < p.L ? p.lse[(long long)z * p.L + i] : INFINITY;
      s_delta[i] = i < p.L ? p.delta[(long long)z * p.L + i] : 0.f;
    }
    const float keep_scale = 1.f / (1.f - p.p_drop);
    uint32_t ka[4][4], va[4][4];
    for (int qb = 0; qb < nqb; ++qb) {
      const int buf = qb & 1;
      if (qb + 1 < nqb) {
        load_tile(smem_u32(tQ) + (buf ^ 1) * TILE_BYTES, seq, pitch, (qb + 1) * BT, p.L, h * HD);
        load_tile(smem_u32(tG) + (buf ^ 1) * TILE_BYTES, gseq, p.D, (qb + 1) * BT, p.L, h * HD);
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      if (qb == 0) {
        load_a_frags(smem_u32(tK), warp * 16, ka);
        load_a_frags(smem_u32(tV), warp * 16, va);
      }
      float st[8][4], dp[8][4];   // S^T and dP^T: [16 keys x 64 queries]
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      mma_a_tileT(st, ka, smem_u32(tQ) + buf * TILE_BYTES);
      mma_a_tileT(dp, va, smem_u32(tG) + buf * TILE_BYTES);
      uint32_t pa[4][4], da[4][4];
      {
        float pd[8][4];   // dropped probabilities (A operand of dV), then dS^T in `st`
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int i = qb * BT + 8 * nt + 2 * cc;      // query columns i, i + 1
          const float2 ls = *reinterpret_cast<const float2*>(s_lse + i);
          const float2 dl = *reinterpret_cast<const float2*>(s_delta + i);
          float p0 = kv0 ? ex2f(fmaf(st[nt][0], p.scale2, -ls.x)) : 0.f;
          float p1 = kv0 ? ex2f(fmaf(st[nt][1], p.scale2, -ls.y)) : 0.f;
          float p2 = kv1 ? ex2f(fmaf(st[nt][2], p.scale2, -ls.x)) : 0.f;
          float p3 = kv1 ? ex2f(fmaf(st[nt][3], p.scale2, -ls.y)) : 0.f;
          float g0 = dp[nt][0], g1 = dp[nt][1], g2 = dp[nt][2], g3 = dp[nt][3];
          float q0 = p0, q1 = p1, q2 = p2, q3 = p3;
          if (p.p_drop > 0.f) {
            const unsigned long long eb = ((unsigned long long)z * p.L + i) * p.L;   // element (query i, key 0)
            const bool c0 = keep_elem(eb + kj0, p.seed, p.p_drop), c1 = keep_elem(eb + p.L + kj0, p.seed, p.p_drop);
            const bool c2 = keep_elem(eb + kj1, p.seed, p.p_drop), c3 = keep_elem(eb + p.L + kj1, p.seed, p.p_drop);
            q0 = c0 ? p0 * keep_scale : 0.f; q1 = c1 ? p1 * keep_scale : 0.f;
            q2 = c2 ? p2 * keep_scale : 0.f; q3 = c3 ? p3 * keep_scale : 0.f;
            g0 = c0 ? g0 * keep_scale : 0.f; g1 = c1 ? g1 * keep_scale : 0.f;
            g2 = c2 ? g2 * keep_scale : 0.f; g3 = c3 ? g3 * keep_scale : 0.f;
          }
          pd[nt][0] = q0; pd[nt][1] = q1; pd[nt][2] = q2; pd[nt][3] = q3;
          st[nt][0] = p0 * (g0 - dl.x); st[nt][1] = p1 * (g1 - dl.y);
          st[nt][2] = p2 * (g2 - dl.x); st[nt][3] = p3 * (g3 - dl.y);
        }
        c_to_a(pd, pa);
        c_to_a(st, da);
      }
      mma_a_tile(dv, pa, smem_u32(tG) + buf * TILE_BYTES);   // dV += P_drop^T dO
      mma_a_tile(dk, da, smem_u32(tQ) + buf * TILE_BYTES);   // dK += dS^T Q
      __syncthreads();
    }
  }
  store_rows(dseq, pitch, k0 + warp * 16, p.L, p.D + h * HD, dk, p.scale, p.scale);
  store_rows(dseq, pitch, k0 + warp * 16, p.L, 2 * p.D + h * HD, dv, 1.f, 1.f);
}

// dQ of one 64-query block: key blocks stream through (padding-only blocks are skipped)
__global__ void __launch_bounds__(128, 3)
bert_attn_bwd_dq_kernel(const BertAttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nkb = (p.L + BT - 1) / BT;
  uint8_t* tQ = smem;
  uint8_t* tG = tQ + TILE_BYTES;
  uint8_t* tK = tG + TILE_BYTES;          // 2 tiles
  uint8_t* tV = tK + 2 * TILE_BYTES;      // 2 tiles
  float* s_mask = reinterpret_cast<float*>(tV + 2 * TILE_BYTES);
  int* s_any = reinterpret_cast<int*>(s_mask + nkb * BT);
  const int z = blockIdx.y, b = z / p.H, h = z - b * p.H;
  const int q0 = blockIdx.x * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, cc = lane & 3;
  const long long pitch = 3LL * p.D;
  const __nv_bfloat16* seq = p.qkv + (long long)b * p.L * pitch;
  const __nv_bfloat16* gseq = p.dout + (long long)b * p.L * p.D;
  load_tile(smem_u32(tQ), seq, pitch, q0, p.L, h * HD);
  load_tile(smem_u32(tG), gseq, p.D, q0, p.L, h * HD);
  cp_async_commit();
  load_key_mask(p, b, s_mask, s_any, nkb);
  int kb = 0;
  while (kb < nkb && !s_any[kb]) ++kb;
  if (kb < nkb) {
    load_tile(smem_u32(tK), seq, pitch, kb * BT, p.L, p.D + h * HD);
    load_tile(smem_u32(tV), seq, pitch, kb * BT, p.L, 2 * p.D + h * HD);
  }
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qa[4][4], ga[4][4];
  load_a_frags(smem_u32(tQ), warp * 16, qa);
  load_a_frags(smem_u32(tG), warp * 16, ga);
  const int qi0 = q0 + warp * 16 + g, qi1 = qi0 + 8;
  const float ls0 = qi0 < p.L ? p.lse[(long long)z * p.L + qi0] : INFINITY;
  const float ls1 = qi1 < p.L ? p.lse[(long long)z * p.L + qi1] : INFINITY;
  const float dl0 = qi0 < p.L ? p.delta[(long long)z * p.L + qi0] : 0.f;
  const float dl1 = qi1 < p.L ? p.delta[(long long)z * p.L + qi1] : 0.f;
  const float keep_scale = 1.f / (1.f - p.p_drop);
  float dq[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
  int buf = 0;
  while (kb < nkb) {
    int nxt = kb + 1;
    while (nxt < nkb && !s_any[nxt]) ++nxt;
    if (nxt < nkb) {
      load_tile(smem_u32(tK) + (buf ^ 1) * TILE_BYTES, seq, pitch, nxt * BT, p.L, p.D + h * HD);
      load_tile(smem_u32(tV) + (buf ^ 1) * TILE_BYTES, seq, pitch, nxt * BT, p.L, 2 * p.D + h * HD);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    mma_a_tileT(s, qa, smem_u32(tK) + buf * TILE_BYTES);
    mma_a_tileT(dp, ga, smem_u32(tV) + buf * TILE_BYTES);
    const unsigned long long e0 = ((unsigned long long)z * p.L + qi0) * p.L + kb * BT;
    const unsigned long long e1 = ((unsigned long long)z * p.L + qi1) * p.L + kb * BT;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int j = 8 * nt + 2 * cc;
      const float2 mk = *reinterpret_cast<const float2*>(s_mask + kb * BT + j);
      const float p0 = ex2f(fmaf(s[nt][0], p.scale2, mk.x) - ls0), p1 = ex2f(fmaf(s[nt][1], p.scale2, mk.y) - ls0);
      const float p2 = ex2f(fmaf(s[nt][2], p.scale2, mk.x) - ls1), p3 = ex2f(fmaf(s[nt][3], p.scale2, mk.y) - ls1);
      float g0 = dp[nt][0], g1 = dp[nt][1], g2 = dp[nt][2], g3 = dp[nt][3];
      if (p.p_drop > 0.f) {
        g0 = keep_elem(e0 + j, p.seed, p.p_drop) ? g0 * keep_scale : 0.f;
        g1 = keep_elem(e0 + j + 1, p.seed, p.p_drop) ? g1 * keep_scale : 0.f;
        g2 = keep_elem(e1 + j, p.seed, p.p_drop) ? g2 * keep_scale : 0.f;
        g3 = keep_elem(e1 + j + 1, p.seed, p.p_drop) ? g3 * keep_scale : 0.f;
      }
      s[nt][0] = p0 * (g0 - dl0); s[nt][1] = p1 * (g1 - dl0);
      s[nt][2] = p2 * (g2 - dl1); s[nt][3] = p3 * (g3 - dl1);
    }
    uint32_t da[4][4];
    c_to_a(s, da);
    mma_a_tile(dq, da, smem_u32(tK) + buf * TILE_BYTES);     // dQ += dS K
    __syncthreads();
    kb = nxt;
    buf ^= 1;
  }
  store_rows(p.dqkv + (long long)b * p.L * pitch, pitch, q0 + warp * 16, p.L, h * HD, dq, p.scale, p.scale);
}

int check(const char* what, int batch, int heads, int seq_len, int hidden, float p_drop) {
  if (batch <= 0 || heads <= 0 || seq_len <= 0) return ctclip::fail(CTCLIP_E_SHAPE, "%s: empty problem", what);
  if (hidden != heads * HD) return ctclip::fail(CTCLIP_E_SHAPE, "%s: head dim must be 64 (hidden %d, heads %d)", what, hidden, heads);
  if (seq_len > 4096) return ctclip::fail(CTCLIP_E_SHAPE, "%s: sequences longer than 4096 are not supported", what);
  if (p_drop < 0.f || p_drop >= 1.f) return ctclip::fail(CTCLIP_E_SHAPE, "%s: dropout must be in [0, 1)", what);
  return ctclip::require_sm100();
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* what) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
  }
  return CTCLIP_OK;
}

}  // namespace

extern "C" int ctclip_bert_attn_fwd(const void* qkv, const long long* mask, int batch, int heads, int seq_len, int hidden,
                                    void* out, float* lse, float p_drop, unsigned seed, void* stream) {
  int rc = check("bert_attn_fwd", batch, heads, seq_len, hidden, p_drop);
  if (rc) return rc;
  if (!qkv || !mask || !out) return ctclip::fail(CTCLIP_E_SHAPE, "bert_attn_fwd: null pointer");
  BertAttnParams p{};
  p.qkv = (const __nv_bfloat16*)qkv; p.mask = mask; p.out = (__nv_bfloat16*)out; p.lse = lse;
  p.B = batch; p.H = heads; p.L = seq_len; p.D = hidden;
  p.scale = 1.f / sqrtf((float)HD); p.scale2 = p.scale * kLog2e; p.p_drop = p_drop; p.seed = seed;
  const int nb = (seq_len + BT - 1) / BT;
  const size_t smem = 5 * TILE_BYTES + (size_t)nb * BT * 4 + (size_t)nb * 4;
  rc = set_smem(bert_attn_fwd_kernel, smem, "bert_attn_fwd");
  if (rc) return rc;
  bert_attn_fwd_kernel<<<dim3(nb, batch * heads), 128, smem, (cudaStream_t)stream>>>(p);
  return ctclip::check_launch("bert_attn_fwd");
}

extern "C" int ctclip_bert_attn_bwd(const void* qkv, const long long* mask, const void* out, const float* lse, const void* dout,
                                    int batch, int heads, int seq_len, int hidden, void* dqkv, float* delta_ws, float p_drop,
                                    unsigned seed, void* stream) {
  int rc = check("bert_attn_bwd", batch, heads, seq_len, hidden, p_drop);
  if (rc) return rc;
  if (!qkv || !mask || !out || !lse || !dout || !dqkv || !delta_ws) return ctclip::fail(CTCLIP_E_SHAPE, "bert_attn_bwd: null pointer");
  BertAttnParams p{};
  p.qkv = (const __nv_bfloat16*)qkv; p.mask = mask; p.out = (__nv_bfloat16*)out; p.lse = const_cast<float*>(lse);
  p.dout = (const __nv_bfloat16*)dout; p.dqkv = (__nv_bfloat16*)dqkv; p.delta = delta_ws;
  p.B = batch; p.H = heads; p.L = seq_len; p.D = hidden;
  p.scale = 1.f / sqrtf((float)HD); p.scale2 = p.scale * kLog2e; p.p_drop = p_drop; p.seed = seed;
  cudaStream_t s = (cudaStream_t)stream;
  const long long rows = (long long)batch * seq_len * heads;
  bert_attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(p);
  rc = ctclip::check_launch("bert_attn_delta");
  if (rc) return rc;
  const int nb = (seq_len + BT - 1) / BT;
  const size_t smem_kv = 6 * TILE_BYTES + (size_t)nb * BT * 8;
  rc = set_smem(bert_attn_bwd_dkv_kernel, smem_kv, "bert_attn_bwd");
  if (rc) return rc;
  bert_attn_bwd_dkv_kernel<<<dim3(nb, batch * heads), 128, smem_kv, s>>>(p);
  rc = ctclip::check_launch("bert_attn_bwd(dkv)");
  if (rc) return rc;
  const size_t smem_q = 6 * TILE_BYTES + (size_t)nb * BT * 4 + (size_t)nb * 4;
  rc = set_smem(bert_attn_bwd_dq_kernel, smem_q, "bert_attn_bwd");
  if (rc) return rc;
  bert_attn_bwd_dq_kernel<<<dim3(nb, batch * heads), 128, smem_q, s>>>(p);
  return ctclip::check_launch("bert_attn_bwd(dq)");
}
