// data_prep volume normalisation on the GPU (reference CTPA_CLIP/data_prep/preprocess_train.py:99-109, resize_array
// :31-42; crop/pad of ct_clip/data.py:155-190):
//   HU = slope*x + intercept (float64) -> clip [-1000,1000] -> /1000 -> float32 -> trilinear resample
//   (align_corners=False) -> optional centre-crop / pad(-1) window into the (240,480,480) training volume.
// One memory-bound kernel: a CTA produces a TD x 1 x TW output tile; the two input rows it needs are staged in
// shared memory with loads that run along the contiguous input axis (z for raw (H,W,N) int16 scans, w for (D,H,W)
// fp32 volumes), the 8-tap gather then reads shared memory only, stores are 256-byte lines along w.
// Bit-exactness: the source index uses a single-rounding fma, and each interpolation level is
// fma(t0, w0, t1*w1) — the exact operation order of ATen's CPU kernel (pinned in oracle/resample_oracle.c).
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {

constexpr int TD = 32, TW = 64, kThreads = 256;

struct PrepParams {
  const void* in;
  float* out;
  int in_is_i16;                       // 1: int16 + HU normalisation, 0: fp32 passthrough
  double slope, intercept;
  int D, H, W;                         // logical input volume (depth, height, width)
  long long sd, sh, sw, sbatch;        // input strides in elements
  int oD, oH, oW;                      // resampled (virtual) grid
  int wd0, wh0, ww0, wdn, whn, wwn;    // window of the resampled grid that is kept (crop)
  int pd0, ph0, pw0;                   // where the window lands in the destination (pad before)
  int tD, tH, tW;                      // destination volume
  long long obatch;
  int kn_max, jn_max;                  // shared tile extents (input depth / width span of one output tile)
  int pre_op, post_op;                 // fp32 input: DataLoader-side arithmetic around the resample (see ctclip_prep_desc)
  float slope32, icpt32;
};

__device__ __forceinline__ void taps(int in, int out, int o, int& i0, int& i1, float& w0, float& w1) {
  if (in == out) { i0 = i1 = o; w0 = 1.f; w1 = 0.f; return; }
  const float scale = __fdiv_rn((float)in, (float)out);
  float src = __fmaf_rn(scale, (float)o + 0.5f, -0.5f);
  src = fmaxf(src, 0.f);
  int a = (int)floorf(src);
  a = min(a, in - 1);
  float l = __fsub_rn(src, (float)a);
  l = fminf(fmaxf(l, 0.f), 1.f);
  i0 = a;
  i1 = a + (a < in - 1 ? 1 : 0);
  w1 = l;
  w0 = __fsub_rn(1.f, l);
}
__device__ __forceinline__ float combine(float t0, float w0, float t1, float w1) {
  return __fmaf_rn(t0, w0, __fmul_rn(t1, w1));
}
__device__ __forceinline__ float load_norm(const PrepParams& p, long long off) {
  if (p.in_is_i16) {
    double v = __dmul_rn(p.slope, (double)reinterpret_cast<const short*>(p.in)[off]);
    v = __dadd_rn(v, p.intercept);
    v = fmin(fmax(v, -1000.0), 1000.0);
    return (float)__ddiv_rn(v, 1000.0);
  }
  return reinterpret_cast<const float*>(p.in)[off];
}

// fp32 input, value arithmetic of the two DataLoaders (one rounding per numpy operation, no contraction):
//   pre_op 2: data_inference.py:81-85   fl(fl(clip(fl(x*1000), -1000, 200) + 400) / 600)
//   pre_op 3: data.py:138               fl(fl(slope*x) + intercept)
__device__ __forceinline__ float pre_f32(const PrepParams& p, float v) {
  if (p.pre_op == 2) {
    v = __fmul_rn(v, 1000.f);
    v = fminf(fmaxf(v, -1000.f), 200.f);
    return __fdiv_rn(__fadd_rn(v, 400.f), 600.f);
  }
  if (p.pre_op == 3) return __fadd_rn(__fmul_rn(p.slope32, v), p.icpt32);
  return v;
}

// exact HU normalisation table: lut[v - lut_lo] = float(clip(slope*v + intercept, -1000, 1000) / 1000) for raw values v in
// [lut_lo, lut_hi]; outside that range the function is saturated, so clamping the index is exact.
__global__ void prep_lut_kernel(float* lut, int lut_lo, int n, double slope, double intercept) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = __dmul_rn(slope, (double)(lut_lo + i));
  v = __dadd_rn(v, intercept);
  v = fmin(fmax(v, -1000.0), 1000.0);
  lut[i] = (float)__ddiv_rn(v, 1000.0);
}

constexpr int OHB = 16;  // output rows (h) produced per CTA: input rows are staged once and reused

// CTA = (TD x OHB x TW) output brick. The input rows (two per output row, shared between consecutive output rows)
// rotate through three shared-memory slots; depth/width taps are computed once per CTA.
__global__ void __launch_bounds__(kThreads)
prep_resample_kernel(const PrepParams p, const float* __restrict__ lut, int lut_lo, int lut_n) {
  extern __shared__ float tile[];  // [3][kn_max][jn_pad] then the LUT copy
  __shared__ int s_d0[TD], s_d1[TD], s_w0[TW], s_w1[TW];
  __shared__ float s_wd0[TD], s_wd1[TD], s_ww0[TW], s_ww1[TW];
  const int jn_pad = p.jn_max | 1;
  const int slot_elems = p.kn_max * jn_pad;
  float* s_lut = tile + 3 * slot_elems;
  const int n_wt = (p.wwn + TW - 1) / TW;
  const int oh_base = p.wh0 + blockIdx.x * OHB;
  const int n_oh = min(OHB, p.wh0 + p.whn - oh_base);
  const int od_base = p.wd0 + (blockIdx.y / n_wt) * TD;
  const int ow_base = p.ww0 + (blockIdx.y % n_wt) * TW;
  const int nd = min(TD, p.wd0 + p.wdn - od_base);
  const int nw = min(TW, p.ww0 + p.wwn - ow_base);
  const long long in_b = (long long)blockIdx.z * p.sbatch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x < TD) {
    int a = 0, b = 0; float x = 0.f, y = 0.f;
    if (threadIdx.x < nd) taps(p.D, p.oD, od_base + threadIdx.x, a, b, x, y);
    s_d0[threadIdx.x] = a; s_d1[threadIdx.x] = b; s_wd0[threadIdx.x] = x; s_wd1[threadIdx.x] = y;
  } else if (threadIdx.x >= 64 && threadIdx.x < 64 + TW) {
    const int t = threadIdx.x - 64;
    int a = 0, b = 0; float x = 0.f, y = 0.f;
    if (t < nw) taps(p.W, p.oW, ow_base + t, a, b, x, y);
    s_w0[t] = a; s_w1[t] = b; s_ww0[t] = x; s_ww1[t] = y;
  }
  for (int i = threadIdx.x; i < lut_n; i += kThreads) s_lut[i] = lut[i];
  __syncthreads();
  const int k_lo = s_d0[0], k_hi = s_d1[nd - 1];
  const int j_lo = s_w0[0], j_hi = s_w1[nw - 1];
  const int kn = k_hi - k_lo + 1, jn = j_hi - j_lo + 1;
  const bool use_lut = lut_n > 0;

  int loaded[3] = {-1, -1, -1};  // input row held by each slot (CTA-uniform bookkeeping)
  auto stage_row = [&](int row) {
    float* dst = tile + (row % 3) * slot_elems;
    const long long row_off = in_b + (long long)row * p.sh;
    if (p.sd == 1) {  // raw scan (H, W, N): depth contiguous -> lanes run along k, warps along j
      for (int j = warp; j < jn; j += kThreads / 32) {
        const long long off = row_off + (long long)(j_lo + j) * p.sw + k_lo;
        for (int k = lane; k < kn; k += 32) {
          float v;
          if (p.in_is_i16) {
            const int raw = reinterpret_cast<const short*>(p.in)[off + k];
            v = use_lut ? s_lut[min(max(raw, lut_lo), lut_lo + lut_n - 1) - lut_lo] : load_norm(p, off + k);
          } else {
            v = pre_f32(p, reinterpret_cast<const float*>(p.in)[off + k]);
          }
          dst[k * jn_pad + j] = v;
        }
      }
    } else {  // (D, H, W): width contiguous -> lanes run along j, warps along k
      for (int k = warp; k < kn; k += kThreads / 32) {
        const long long off = row_off + (long long)(k_lo + k) * p.sd + (long long)j_lo * p.sw;
        for (int j = lane; j < jn; j += 32) {
          float v;
          if (p.in_is_i16) {
            const int raw = reinterpret_cast<const short*>(p.in)[off + (long long)j * p.sw];
            v = use_lut ? s_lut[min(max(raw, lut_lo), lut_lo + lut_n - 1) - lut_lo] : load_norm(p, off + (long long)j * p.sw);
          } else {
            v = pre_f32(p, reinterpret_cast<const float*>(p.in)[off + (long long)j * p.sw]);
          }
          dst[k * jn_pad + j] = v;
        }
      }
    }
  };

  float* out_b = p.out + (long long)blockIdx.z * p.obatch;
  for (int t = 0; t < n_oh; ++t) {
    const int oh = oh_base + t;
    int h0, h1; float wh0, wh1;
    taps(p.H, p.oH, oh, h0, h1, wh0, wh1);
    bool staged = false;
    if (loaded[h0 % 3] != h0) { stage_row(h0); loaded[h0 % 3] = h0; staged = true; }
    if (loaded[h1 % 3] != h1) { stage_row(h1); loaded[h1 % 3] = h1; staged = true; }
    if (staged) __syncthreads();
    const float* r0 = tile + (h0 % 3) * slot_elems;
    const float* r1 = tile + (h1 % 3) * slot_elems;
    const int dh = oh - p.wh0 + p.ph0;
    for (int e = threadIdx.x; e < nd * TW; e += kThreads) {
      const int x = e % TW, z = e / TW;
      if (x >= nw) continue;
      const int ka = (s_d0[z] - k_lo) * jn_pad, kb = (s_d1[z] - k_lo) * jn_pad;
      const int ja = s_w0[x] - j_lo, jb = s_w1[x] - j_lo;
      const float wa = s_ww0[x], wb = s_ww1[x];
      const float a = combine(r0[ka + ja], wa, r0[ka + jb], wb);   // d0, h0
      const float b = combine(r1[ka + ja], wa, r1[ka + jb], wb);   // d0, h1
      const float c = combine(r0[kb + ja], wa, r0[kb + jb], wb);   // d1, h0
      const float d = combine(r1[kb + ja], wa, r1[kb + jb], wb);   // d1, h1
      const float ab = combine(a, wh0, b, wh1);
      const float cd = combine(c, wh0, d, wh1);
      float v = combine(ab, s_wd0[z], cd, s_wd1[z]);
      if (p.post_op == 1) v = __fdiv_rn(fminf(fmaxf(v, -1000.f), 1000.f), 1000.f);   // data.py:150-152
      const int dd = od_base + z - p.wd0 + p.pd0;
      const int dw = ow_base + x - p.ww0 + p.pw0;
      out_b[((long long)dd * p.tH + dh) * p.tW + dw] = v;
    }
    // the next output row may overwrite the slot of a row that is no longer needed: wait for all readers
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------ fast path
// Raw scans as NIfTI stores them — int16 (H, W, N), depth contiguous, N % 8 == 0 — resampled into (D', H', W') fp32.
// A CTA owns an (ftd x FOH x fw) output brick: lane = output column, warp = a run of consecutive output depths. The
// input rows (h) the brick needs rotate through three shared-memory slots [k][j] (fp32, HU-normalised while staging;
// global loads are 16-byte chunks along the contiguous depth axis, prefetched into registers one output row ahead).
//
// Register marching, two levels: a thread keeps the w-interpolated values W(k) of its <= KMAX input depth planes for
// the two input rows (h0, h1) of the current output row IN REGISTERS. Stepping to the next output row re-uses the h1
// row as the new h0 (one new row = 2 LDS + 2 FP per plane); P(k) = lerp_h(W_h0(k), W_h1(k)) is then 2 FP per plane and
// every output is one more lerp of P(k), P(k+1) — ~18 instructions per output instead of 8 LDS + 8 index reads + 14 FP.
// The w taps live in registers, the d / h taps in small shared tables; fw <= 32 is chosen by the host so that the input
// columns of one warp span at most 32 words (conflict-free). Same operation order as the generic kernel (w, then h,
// then d, each level fma(t0, w0, t1 * w1)) -> bit-identical results.
//
// HU normalisation: ARITH = true when slope == 1 and the intercept is an integer (the usual CT rescale): then
// c = clip(raw + intercept) is an integer in [-1000, 1000] and float(double(c) / 1000.0) == q1 with
// q0 = c * 0.001f, q1 = fma(fma(-q0, 1000, c), 0.001f, q0) for all 2001 values (checked exhaustively in
// tests/test_resample_oracle.py) — three FP32 instructions instead of a random-bank table look-up. Otherwise the exact
// LUT is used.
constexpr int FG = 8, FOH = 96;            // depth groups (warps), oh per CTA
// Variant NC = 1: lane = one output column, 16 register planes (12 outputs per warp run at 4:3), row pitch 33 words.
// Variant NC = 2: lane = TWO output columns `fw` apart (each group of fw columns spans <= 32 input words, so every
// shared-memory load stays conflict-free); all three lerp levels run as packed FMUL2 / FFMA2 on the column pair, the depth /
// height tap loads, the emission tests and the per-row bookkeeping are shared by the pair — about 0.6x the instructions
// per output voxel. 10 register planes (6 outputs per warp run at 4:3), row pitch 69 words, 3 chunks per thread per row.
template <int NC> struct FastCfg;
template <> struct FastCfg<1> { static constexpr int KMAX = 16, CPR = 2, JP = 33; using W = float; };
template <> struct FastCfg<2> { static constexpr int KMAX = 10, CPR = 3, JP = 69; using W = float2; };

__device__ __forceinline__ float lerpw(float t0, float w0, float t1, float w1) { return combine(t0, w0, t1, w1); }
__device__ __forceinline__ float2 lerpw(float2 t0, float2 w0, float2 t1, float2 w1) { return ptx::ffma2(t0, w0, ptx::fmul2(t1, w1)); }
__device__ __forceinline__ void splat(float& d, float v) { d = v; }
__device__ __forceinline__ void splat(float2& d, float v) { d = make_float2(v, v); }

template <bool ARITH, int NC>
__global__ void __launch_bounds__(256, 2)
prep_hwn_i16_kernel(const PrepParams p, const float* __restrict__ lut, int lut_lo, int lut_n, int icpt, int ftd, int fw,
                    int kn8_max) {
  using Cfg = FastCfg<NC>;
  using W = typename Cfg::W;
  constexpr int FKMAX = Cfg::KMAX, FCPR = Cfg::CPR, FJP = Cfg::JP;
  extern __shared__ float tile[];  // [3][kn8_max + 1][FJP] + FKMAX rows of slack | d taps [ftd] | h taps [FOH] | lut
  const int slot_elems = (kn8_max + 1) * FJP;
  float4* s_dtap = reinterpret_cast<float4*>(tile + ((3 * slot_elems + FKMAX * FJP + 3) & ~3));
  float4* s_htap = s_dtap + ftd;
  float* s_lut = reinterpret_cast<float*>(s_htap + FOH);
  const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5;
  const int ctw = NC * fw;   // output columns per CTA
  const int n_wt = (p.wwn + ctw - 1) / ctw;
  const int oh_base = p.wh0 + blockIdx.x * FOH;
  const int n_oh = min(FOH, p.wh0 + p.whn - oh_base);
  const int od_base = p.wd0 + (blockIdx.y / n_wt) * ftd;
  const int ow_base = p.ww0 + (blockIdx.y % n_wt) * ctw;
  const int nd = min(ftd, p.wd0 + p.wdn - od_base);
  const int nw = min(ctw, p.ww0 + p.wwn - ow_base);
  const short* in16 = reinterpret_cast<const short*>(p.in) + (long long)blockIdx.z * p.sbatch;

  for (int z = tid; z < nd; z += 256) {
    int a, b; float x, y;
    taps(p.D, p.oD, od_base + z, a, b, x, y);
    s_dtap[z] = make_float4(__int_as_float(a), x, y, __int_as_float(b));
  }
  if (tid < n_oh) {
    int a, b; float x, y;
    taps(p.H, p.oH, oh_base + tid, a, b, x, y);
    s_htap[tid] = make_float4(__int_as_float(a), x, y, __int_as_float(b));
  }
  if (!ARITH)
    for (int i = tid; i < lut_n; i += 256) s_lut[i] = lut[i];
  // CTA-uniform input extents (every thread evaluates the same taps)
  int k_lo, k_hi, j_lo, j_hi, t0, t1; float f0, f1;
  taps(p.D, p.oD, od_base, k_lo, t1, f0, f1);
  taps(p.D, p.oD, od_base + nd - 1, t0, t1, f0, f1);
  k_hi = min(p.D - 1, max(t1, t0 + 1));
  taps(p.W, p.oW, ow_base, j_lo, t1, f0, f1);
  taps(p.W, p.oW, ow_base + nw - 1, t0, j_hi, f0, f1);
  const int k_lo8 = k_lo & ~7;
  const int kc_n = (k_hi - k_lo8) / 8 + 1;   // 16-byte chunks per input column
  const int jn = j_hi - j_lo + 1;
  const int chunks = kc_n * jn;
  // this thread's w taps: column lane + c * fw of the CTA's column tile
  bool ow_ok[NC];
  int ja[NC], jb[NC];
  W wa, wb;
  {
    float fa[NC], fb[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = lane + c * fw;
      ow_ok[c] = lane < fw && col < nw;
      taps(p.W, p.oW, ow_base + (ow_ok[c] ? col : nw - 1), ja[c], jb[c], fa[c], fb[c]);
      ja[c] -= j_lo; jb[c] -= j_lo;
    }
    if constexpr (NC == 1) { wa = fa[0]; wb = fb[0]; } else { wa = make_float2(fa[0], fa[1]); wb = make_float2(fb[0], fb[1]); }
  }
  const int lut_hi = lut_lo + lut_n - 1;
  // chunk -> (column j, depth chunk kc) of this thread's staging slots (same for every row)
  int cj[FCPR], ck[FCPR];
#pragma unroll
  for (int u = 0; u < FCPR; ++u) {
    const int idx = tid + u * 256;
    ck[u] = idx / jn;
    cj[u] = idx - ck[u] * jn;
  }

  const short* cptr[FCPR];   // chunk address inside input row 0
#pragma unroll
  for (int u = 0; u < FCPR; ++u) cptr[u] = in16 + (long long)(j_lo + cj[u]) * p.sw + k_lo8 + 8 * ck[u];
  const bool last_chunk_dup = (k_lo8 + 8 * kc_n == p.D);   // the brick reaches the last plane: duplicate it at k = D
  auto row_load = [&](int row, uint4 (&pf)[FCPR]) {
    const long long ro = (long long)row * p.sh;
#pragma unroll
    for (int u = 0; u < FCPR; ++u)
      if (tid + u * 256 < chunks) pf[u] = __ldg(reinterpret_cast<const uint4*>(cptr[u] + ro));
  };
  auto row_store = [&](int row, const uint4 (&pf)[FCPR]) {
    float* dst = tile + (row % 3) * slot_elems;
#pragma unroll
    for (int u = 0; u < FCPR; ++u) {
      if (tid + u * 256 < chunks) {
        const uint32_t w4[4] = {pf[u].x, pf[u].y, pf[u].z, pf[u].w};
        float* d = dst + (8 * ck[u]) * FJP + cj[u];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r0 = (int)(short)(w4[e] & 0xffffu), r1 = (int)(short)(w4[e] >> 16);
          float2 y;
          if (ARITH) {
            const float2 cf = make_float2((float)min(max(r0 + icpt, -1000), 1000), (float)min(max(r1 + icpt, -1000), 1000));
            const float2 k3 = make_float2(0.001f, 0.001f);
            const float2 q0 = ptx::fmul2(cf, k3);
            y = ptx::ffma2(ptx::ffma2(make_float2(-q0.x, -q0.y), make_float2(1000.f, 1000.f), cf), k3, q0);
          } else {
            y.x = s_lut[min(max(r0, lut_lo), lut_hi) - lut_lo];
            y.y = s_lut[min(max(r1, lut_lo), lut_hi) - lut_lo];
          }
          d[(2 * e) * FJP] = y.x;
          d[(2 * e + 1) * FJP] = y.y;
          if (e == 3 && last_chunk_dup && ck[u] == kc_n - 1) d[8 * FJP] = y.y;   // P(D) := P(D-1): no select in the march
        }
      }
    }
  };

  __syncthreads();  // LUT and taps visible
  // this thread's run of output depths [zbeg, zend) and the input planes kf .. kf + kn_t - 1 it needs
  const int opt = ftd / FG;
  const int zbeg = grp * opt, zend = min(zbeg + opt, nd);
  int kf = 0, kn_t = 0;
  if (zbeg < zend) {
    kf = __float_as_int(s_dtap[zbeg].x);
    kn_t = min(__float_as_int(s_dtap[zend - 1].x) + 1, p.D - 1) - kf + 1;
  }
  // depth down-sampling (host-checked): at most one output has its lower tap on a given plane -> a 16-bit mask per warp
  unsigned emask = 0;
  for (int z = zbeg; z < zend; ++z) emask |= 1u << (__float_as_int(s_dtap[z].x) - kf);
  int ld0 = -1, ld1 = -1, ld2 = -1;  // input row held by each slot
  auto held = [&](int row) { const int m = row % 3; return (m == 0 ? ld0 : m == 1 ? ld1 : ld2) == row; };
  auto hold = [&](int row) { const int m = row % 3; if (m == 0) ld0 = row; else if (m == 1) ld1 = row; else ld2 = row; };
  uint4 pfa[FCPR], pfb[FCPR];
  {
    const float4 ht = s_htap[0];
    const int h0 = __float_as_int(ht.x), h1 = __float_as_int(ht.w);
    row_load(h0, pfa);
    if (h1 != h0) row_load(h1, pfb);
    row_store(h0, pfa); hold(h0);
    if (h1 != h0) { row_store(h1, pfb); hold(h1); }
  }
  __syncthreads();

  int woff[NC], djb[NC];                         // word offset of W(kf) inside a slot ; jb - ja (0 or 1)
#pragma unroll
  for (int c = 0; c < NC; ++c) { woff[c] = (kf - k_lo8) * FJP + ja[c]; djb[c] = jb[c] - ja[c]; }
  W wA[FKMAX], wB[FKMAX];                        // w-interpolated planes of two input rows (rowA, rowB)
  int rowA = -1, rowB = -1;
  auto wrow = [&](int row, W (&w)[FKMAX]) {      // planes beyond kn_t read slack / stale words: never used
    const float* ra = tile + (row % 3) * slot_elems + woff[0];
    const float* rb = ra + djb[0];
    if constexpr (NC == 1) {
#pragma unroll
      for (int kk = 0; kk < FKMAX; ++kk) w[kk] = lerpw(ra[kk * FJP], wa, rb[kk * FJP], wb);
    } else {
      const float* ra1 = tile + (row % 3) * slot_elems + woff[NC - 1];
      const float* rb1 = ra1 + djb[NC - 1];
#pragma unroll
      for (int kk = 0; kk < FKMAX; ++kk)
        w[kk] = lerpw(make_float2(ra[kk * FJP], ra1[kk * FJP]), wa, make_float2(rb[kk * FJP], rb1[kk * FJP]), wb);
    }
  };
  const long long oplane = (long long)p.tH * p.tW;
  float* out_t = p.out + (long long)blockIdx.z * p.obatch + (long long)(od_base + zbeg - p.wd0 + p.pd0) * oplane +
                 (ow_base + lane - p.ww0 + p.pw0);
  bool synced_prev = true;   // the barrier after the initial row stores
  for (int t = 0; t < n_oh; ++t) {
    const int oh = oh_base + t;
    const float4 ht = s_htap[t];
    const int h0 = __float_as_int(ht.x), h1 = __float_as_int(ht.w);
    const float wh0 = ht.y, wh1 = ht.z;
    int na = -1, nb = -1;
    if (t + 1 < n_oh) {
      const float4 hx = s_htap[t + 1];
      const int g0 = __float_as_int(hx.x), g1 = __float_as_int(hx.w);
      if (!held(g0)) na = g0;
      if (g1 != g0 && !held(g1)) nb = g1;
      if (na >= 0) row_load(na, pfa);
      if (nb >= 0) row_load(nb, pfb);
    }
    if (kn_t > 0) {
      float* optr = out_t + (long long)(oh - p.wh0 + p.ph0) * p.tW;
      W vh0, vh1;
      splat(vh0, wh0);
      splat(vh1, wh1);
      auto march = [&](const W (&x)[FKMAX], const W (&y)[FKMAX]) {   // x: row h0, y: row h1
        W pc = lerpw(x[0], vh0, y[0], vh1);
        const float4* wp = s_dtap + zbeg;     // depth taps of the next output of this run
        float* op = optr;
#pragma unroll
        for (int kk = 0; kk < FKMAX; ++kk) {
          const W pn = (kk + 1 < FKMAX) ? lerpw(x[(kk + 1) % FKMAX], vh0, y[(kk + 1) % FKMAX], vh1) : pc;
          if (emask & (1u << kk)) {           // warp-uniform
            const float4 tp = *wp++;
            W t0, t1;
            splat(t0, tp.y);
            splat(t1, tp.z);
            const W v = lerpw(pc, t0, pn, t1);
            if constexpr (NC == 1) {
              if (ow_ok[0]) *op = v;
            } else {
              if (ow_ok[0]) *op = v.x;
              if (ow_ok[NC - 1]) op[fw] = v.y;
            }
            op += oplane;
          }
          pc = pn;
        }
      };
      // the two register planes swap roles instead of being copied: whichever already holds h0 stays
      if (rowA == h0) {
        if (h1 != h0 && rowB != h1) { wrow(h1, wB); rowB = h1; }
        if (h1 != h0) march(wA, wB); else march(wA, wA);
      } else if (rowB == h0) {
        if (h1 != h0 && rowA != h1) { wrow(h1, wA); rowA = h1; }
        if (h1 != h0) march(wB, wA); else march(wB, wB);
      } else {
        wrow(h0, wA); rowA = h0;
        if (h1 != h0 && rowB != h1) { wrow(h1, wB); rowB = h1; }
        if (h1 != h0) march(wA, wB); else march(wA, wA);
      }
    }
    if (na >= 0 || nb >= 0) {
      // Slot reads happen only in wrow() at the top of an iteration and only for rows h0 / h1. A new row whose slot is
      // neither of those two was last read in an EARLIER iteration; if that iteration ended with the barrier below, every
      // warp is past those reads and the store needs no barrier of its own (the usual 512 -> 480 step: rows r, r+1 in use,
      // row r+2 replaces r-1). Otherwise (two new rows, one of them landing on h0's slot; or no barrier last time) wait.
      const int s0 = h0 % 3, s1 = h1 % 3;
      const bool clash = (na >= 0 && (na % 3 == s0 || na % 3 == s1)) || (nb >= 0 && (nb % 3 == s0 || nb % 3 == s1));
      if (clash || !synced_prev) __syncthreads();  // every reader of the slots being replaced is done
      if (na >= 0) { row_store(na, pfa); hold(na); }
      if (nb >= 0) { row_store(nb, pfb); hold(nb); }
      __syncthreads();
      synced_prev = true;
    } else {
      synced_prev = false;
    }
  }
}

// ------------------------------------------------------------------------------------------------ fast path, v2
// Same brick decomposition and the same register marching as prep_hwn_i16_kernel, re-laid-out around 128-bit shared-memory
// traffic and packed arithmetic (the v1 kernel issued ~35 instructions per output voxel, a third of them scalar LDS / STS):
//  * slots are [j][k] — the depth axis k is contiguous, pitch KP = kn8 + 4 floats (KP / 4 odd: the 16-byte accesses of
//    eight consecutive columns fall into eight different bank groups). A staging thread converts its 16-byte chunk (8 raw
//    int16 along k) and writes it back with two STS.128; the w-interpolation of a thread reads its 8 planes of a column
//    with two LDS.128 per tap instead of eight scalar loads.
//  * HU normalisation on packed 16-bit integers: clip in the raw domain (VIMNMX.S16x2), one packed add of (intercept + 1024),
//    int -> float through the 2^23 mantissa trick (PRMT + one packed FADD, no I2F), then the same exact 3-instruction
//    Newton division as v1 on FMUL2 / FFMA2.
//  * all three lerp levels are packed along the DEPTH axis (adjacent planes of one column sit in adjacent registers), the
//    d taps of a thread's run live in registers, and for the production 4:3 depth ratio (three outputs per four input
//    planes, no plane shared between quads) the emission is straight-line code.
// A warp owns 6 output depths = an aligned window of 8 input planes; lane = output column (fw <= 32 per CTA, chosen by the
// host so that the CTA's input columns span <= 32). Same operation order as the generic kernel -> bit-identical results.
constexpr int V2_OPT = 6, V2_KW = 8, V2_CPR = 2;
constexpr float kMagic = 8388608.f + 1024.f;   // float(0x4B000000 | u) - kMagic == u - 1024 exactly for 0 <= u < 2^16

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ float4 lds128(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Input-row pipeline of a CTA (rows are consumed in increasing order without gaps — host-checked: H <= 2 oH):
//   global --cp.async (LDGSTS), RING rows ahead--> raw ring --convert--> fp32 slots (4; row r lives in slot r & 3) --> wrow
// Each thread copies, and later reads back, its OWN 16-byte chunks of the raw ring, so the ring needs neither a barrier nor
// a swizzle; RING - 1 rows (3 x 4 KB per CTA, ~40 KB per SM) stay in flight — the register-prefetch version kept ~6 KB per
// SM in flight against the ~11 KB that 2.0 TB/s of loads x ~800 ns ask for. With FOUR fp32 slots a new row never lands on a
// slot the current output row still reads (rows h0, h0+1 in use; h0+2 and h0+3 go to the two other slots), so ONE barrier
// per converted row — after the store — is enough. All slot / ring / global addresses advance incrementally.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  // .ca (through L1): the lanes of a warp that share a 32-byte sector are served by ONE sector fetch; the .cg (L1 bypass)
  // form fetched a sector per lane — 2x the L2 -> SM traffic for 16-byte pieces (ncu: 32 sectors per request)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }

constexpr int V2_RING = 4, V2_SLOTS = 4;

// NC: output columns per lane (lane, lane + fw): the per-row bookkeeping (h taps, row pipeline, barrier, branches — more than
// a third of the NC = 1 instruction stream) is shared by the column pair. MINB: resident CTAs per SM the register cap allows.
// TMA: true (CTCLIP_PREP_V2_TMA=1, when the brick's raw rows fit a 64-plane x 32 * NC-column box) = ONE thread fetches each raw row with a
// 4-D `cp.async.bulk.tensor` (SWIZZLE_128B, one 128-byte line per input column) into the ring, completion on an mbarrier per
// slot, refill right after the CTA barrier that follows the row's conversion; false = every thread copies its own 16-byte
// chunks with cp.async (LDGSTS). ncu on the LDGSTS ring: ~18 LSU wavefronts per 512-byte warp request (one per returning
// sector) in a kernel whose LSU data pipe is its busiest unit — TMA writes shared memory without the LSU.
// A register-prefetch pipeline (LDG.128 three rows ahead) was measured too: 4.7 ms against 3.05 ms per 32 scans (the register
// shifts of the pipeline wait for the loads they move). CPR: 16-byte chunks per thread and row (1 | 2).
template <int NC, int MINB, bool TMA, int CPR>
__global__ void __launch_bounds__(256, MINB)
prep_hwn_i16_v2_kernel(const __grid_constant__ CUtensorMap tmap, const PrepParams p, int icpt, int fw, int kn8, int jn_max,
                       int ring_chunks) {
  extern __shared__ __align__(16) float tile[];  // [4][jn_max][KP] + slack | d taps [48] | h taps [FOH + 1] | raw ring [4][ring_chunks] x 16 B
  const int KP = kn8 + 4;
  const int slot_elems = jn_max * KP;
  float4* s_dtap = reinterpret_cast<float4*>(tile + V2_SLOTS * slot_elems + 16);
  float4* s_htap = s_dtap + FG * V2_OPT;
  uint4* s_ring = reinterpret_cast<uint4*>(s_htap + FOH + 1);
  const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5;
  constexpr int ftd = FG * V2_OPT;
  const int ctw = NC * fw;                       // output columns per CTA
  const int n_wt = (p.wwn + ctw - 1) / ctw;
  const int oh_base = p.wh0 + blockIdx.x * FOH;
  const int n_oh = min(FOH, p.wh0 + p.whn - oh_base);
  const int od_base = p.wd0 + (blockIdx.y / n_wt) * ftd;
  const int ow_base = p.ww0 + (blockIdx.y % n_wt) * ctw;
  const int nd = min(ftd, p.wd0 + p.wdn - od_base);
  const int nw = min(ctw, p.ww0 + p.wwn - ow_base);
  const short* in16 = reinterpret_cast<const short*>(p.in) + (long long)blockIdx.z * p.sbatch;

  if (tid < nd) {
    int a, b; float x, y;
    taps(p.D, p.oD, od_base + tid, a, b, x, y);
    s_dtap[tid] = make_float4(x, y, __int_as_float(a), __int_as_float(b));   // (w0, w1, k0, k1)
  }
  if (tid <= n_oh) {   // one entry past the end (a copy of the last row): the loop prefetches the NEXT row's taps
    int a, b; float x, y;
    taps(p.H, p.oH, oh_base + min(tid, n_oh - 1), a, b, x, y);
    s_htap[tid] = make_float4(__int_as_float(a), x, y, __int_as_float(b));
  }
  int k_lo, k_hi, j_lo, j_hi, t0, t1; float f0, f1;
  taps(p.D, p.oD, od_base, k_lo, t1, f0, f1);
  taps(p.D, p.oD, od_base + nd - 1, t0, t1, f0, f1);
  k_hi = min(p.D - 1, max(t1, t0 + 1));
  taps(p.W, p.oW, ow_base, j_lo, t1, f0, f1);
  taps(p.W, p.oW, ow_base + nw - 1, t0, j_hi, f0, f1);
  const int k_lo8 = k_lo & ~7;
  const int kc_n = (k_hi - k_lo8) / 8 + 1;   // 16-byte chunks per input column
  const int jn = j_hi - j_lo + 1;
  const int chunks = kc_n * jn;
  int h_first, h_last;
  taps(p.H, p.oH, oh_base, h_first, t1, f0, f1);
  taps(p.H, p.oH, oh_base + n_oh - 1, t0, h_last, f0, f1);

  // this thread's staging chunks: global address in the row being prefetched, byte offset inside an fp32 slot
  const bool last_chunk_dup = (k_lo8 + 8 * kc_n == p.D);   // the brick reaches the last plane: duplicate it at k = D
  const short* gsrc[CPR];
  uint32_t cdst[CPR], csrc[CPR];
  bool cuse[CPR], cdup[CPR];
  const uint32_t tile_u32 = ptx::smem_u32(tile);
  // chunk -> (column cj, depth chunk ck). Lanes run along the CONTIGUOUS depth axis: with 8 chunks per column a quarter-warp
  // takes chunks 0-3 (or 4-7) of two adjacent columns, so a warp's cp.async reads 4 columns x 128 contiguous bytes (16 full
  // sectors; one lane per column read 16 of every 32-byte sector and cost 32 LSU wavefronts per instruction), and its two
  // STS.128 per chunk fall into eight different bank groups (column pitch KP = 17 x 16 B).
#pragma unroll
  for (int u = 0; u < CPR; ++u) {
    const int idx = tid + u * 256;
    int ck, cj;
    if (kc_n == 8) {
      const int r = idx & 15;
      cj = 2 * (idx >> 4) + ((r >> 2) & 1);
      ck = (r & 3) + 4 * (r >> 3);
    } else {
      cj = idx / kc_n;
      ck = idx - cj * kc_n;
    }
    cuse[u] = cj < jn;
    cdup[u] = cuse[u] && last_chunk_dup && ck == kc_n - 1;
    gsrc[u] = in16 + (long long)h_first * p.sh + (long long)(j_lo + cj) * p.sw + k_lo8 + 8 * ck;
    cdst[u] = tile_u32 + (uint32_t)(cj * KP + 8 * ck) * 4u;
    csrc[u] = TMA ? (uint32_t)(cj * 128 + ((ck ^ (cj & 7)) << 4)) : (uint32_t)u * 4096u;   // chunk offset inside a ring slot
  }
  const uint32_t slot_bytes = (uint32_t)slot_elems * 4u, ring_bytes = (uint32_t)ring_chunks * 16u;   // ring_bytes: 4096 or 8192
  const uint32_t ring_mask = V2_RING * ring_bytes - 1;
  // TMA ring: 1024-byte aligned (128-byte swizzle atom), row of column cj at cj * 128, chunk ck at ((ck ^ (cj & 7)) << 4)
  const uint32_t ring_al = (ptx::smem_u32(s_ring) + 1023u) & ~1023u;
  uint8_t* ring_ptr = reinterpret_cast<uint8_t*>(tile) + (ring_al - ptx::smem_u32(tile));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring_ptr + V2_RING * ring_bytes);
  const uint32_t ring_u32 = TMA ? ring_al : ptx::smem_u32(s_ring) + tid * 16;
  uint32_t ring_wr = 0, ring_rd = 0;            // byte offsets of the slots of the next row to prefetch / to take
  int pf_left = h_last - h_first + 1;           // (cp.async) rows still to prefetch
  int pf_row = h_first;                         // (TMA) next row to fetch
  uint32_t rd_parity = 0;                       // (TMA) phase of the slot being taken
  int conv_hi = h_first - 1;                    // highest row converted so far
  uint32_t cv_slot = (uint32_t)(h_first & (V2_SLOTS - 1)) * slot_bytes;   // fp32 slot of row conv_hi + 1
  auto prefetch_next = [&]() {                  // cp.async flavour: every thread, right after it consumed its chunks of a row
    if (pf_left > 0) {
#pragma unroll
      for (int u = 0; u < CPR; ++u)
        if (cuse[u]) { cp_async16(ring_u32 + ring_wr + csrc[u], gsrc[u]); gsrc[u] += p.sh; }
    }
    cp_async_commit();                          // one group per row, empty beyond the brick's last row
    ring_wr = (ring_wr + ring_bytes) & ring_mask;
    --pf_left;
  };
  auto tma_refill = [&]() {                     // TMA flavour: thread 0, after the barrier that ends the reads of the freed slots
    while (pf_row <= h_last && pf_row <= conv_hi + V2_RING) {
      uint64_t* bar = full_bar + (ring_wr / ring_bytes);
      ptx::mbar_expect_tx(bar, ring_bytes);
      ptx::tma_load_4d(ring_ptr + ring_wr, &tmap, bar, k_lo8, j_lo, pf_row, (int)blockIdx.z);
      ring_wr = (ring_wr + ring_bytes) & ring_mask;
      ++pf_row;
    }
  };
  // raw clip bounds and bias as packed int16 pairs: c + 1024 = clamp(raw, -1000 - icpt, 1000 - icpt) + (icpt + 1024)
  const uint32_t lo2 = (uint32_t)(uint16_t)(short)(-1000 - icpt) * 0x10001u, hi2 = (uint32_t)(uint16_t)(short)(1000 - icpt) * 0x10001u;
  const uint32_t bias2 = (uint32_t)(uint16_t)(short)(icpt + 1024) * 0x10001u;
  auto convert_next = [&]() {                   // raw row conv_hi + 1: ring -> HU-normalised fp32 slot; refill its ring slot
    if constexpr (TMA) ptx::mbar_wait(full_bar + (ring_rd / ring_bytes), rd_parity); else cp_async_wait<V2_RING - 1>();
#pragma unroll
    for (int u = 0; u < CPR; ++u) {
      if (cuse[u]) {
        const uint4 pf = lds128u(ring_u32 + ring_rd + csrc[u]);
        const uint32_t w4[4] = {pf.x, pf.y, pf.z, pf.w};
        float2 y[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t c2 = __vadd2(__vmaxs2(__vmins2(w4[e], hi2), lo2), bias2);
          const float2 m = make_float2(__uint_as_float(prmt(c2, 0x4B000000u, 0x7610u)),
                                       __uint_as_float(prmt(c2, 0x4B000000u, 0x7632u)));
          const float2 cf = ptx::fadd2(m, make_float2(-kMagic, -kMagic));          // exact: clip(raw + icpt) as float
          const float2 k3 = make_float2(0.001f, 0.001f);
          const float2 q0 = ptx::fmul2(cf, k3);
          y[e] = ptx::ffma2(ptx::ffma2(q0, make_float2(-1000.f, -1000.f), cf), k3, q0);   // == float(double(c) / 1000.0)
        }
        const uint32_t d = cdst[u] + cv_slot;
        sts128(d, y[0].x, y[0].y, y[1].x, y[1].y);
        sts128(d + 16, y[2].x, y[2].y, y[3].x, y[3].y);
        if (cdup[u]) sts32(d + 32, y[3].y);     // P(D) := P(D-1): no select in the march
      }
    }
    ++conv_hi;
    cv_slot += slot_bytes;
    if (cv_slot == V2_SLOTS * slot_bytes) cv_slot = 0;
    ring_rd = (ring_rd + ring_bytes) & ring_mask;
    if constexpr (TMA) { if (ring_rd == 0) rd_parity ^= 1u; } else { prefetch_next(); }
  };

  if constexpr (TMA) {
    if (tid == 0) {
      ptx::prefetch_tmap(&tmap);
#pragma unroll
      for (int r = 0; r < V2_RING; ++r) ptx::mbar_init(full_bar + r, 1);
      ptx::fence_barrier_init();
    }
  }
  __syncthreads();  // taps (and the ring barriers) visible
  // this warp's run of output depths [zbeg, zend), its aligned window of V2_KW input planes starting at kw0, its d taps
  const int zbeg = grp * V2_OPT, zend = min(zbeg + V2_OPT, nd);
  int kw0 = 0;
  unsigned emask = 0;
  constexpr bool kTapsInRegs = (NC == 1 && MINB == 2);   // elsewhere registers are tight: the d taps are broadcast LDS.64 per output row
  float2 dt[V2_OPT];
#pragma unroll
  for (int i = 0; i < V2_OPT; ++i) dt[i] = make_float2(0.f, 0.f);
  if (zbeg < zend) {
    kw0 = __float_as_int(s_dtap[zbeg].z) & ~3;
#pragma unroll
    for (int i = 0; i < V2_OPT; ++i)
      if (zbeg + i < zend) {
        const float4 tp = s_dtap[zbeg + i];
        emask |= 1u << (__float_as_int(tp.z) - kw0);
        if (kTapsInRegs) dt[i] = make_float2(tp.x, tp.y);
      }
  }
  const float4* dtp = s_dtap + zbeg;
  const bool quad43 = (emask == 0x77u);   // 4:3 depth ratio: outputs on planes 0,1,2 and 4,5,6 of the window
  // this thread's columns and their w taps
  bool ow_ok[NC];
  uint32_t woff_a[NC], woff_b[NC];
  float2 wa[NC], wb[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = lane + c * fw;
    ow_ok[c] = lane < fw && col < nw;
    int ja, jb; float fa, fb;
    taps(p.W, p.oW, ow_base + (ow_ok[c] ? col : nw - 1), ja, jb, fa, fb);
    woff_a[c] = tile_u32 + (uint32_t)((ja - j_lo) * KP + (kw0 - k_lo8)) * 4u;
    woff_b[c] = tile_u32 + (uint32_t)((jb - j_lo) * KP + (kw0 - k_lo8)) * 4u;
    wa[c] = make_float2(fa, fa); wb[c] = make_float2(fb, fb);
  }
  const bool active = zbeg < zend && ow_ok[0];   // lanes beyond the tile's columns / warps beyond its depths only stage

  if constexpr (TMA) {
    if (tid == 0) tma_refill();                 // rows h_first .. h_first + 3 (barriers were initialised before the tap barrier)
  } else {
#pragma unroll
    for (int r = 0; r < V2_RING; ++r) prefetch_next();
  }
  {
    const int h1_0 = __float_as_int(s_htap[0].w);
    do { convert_next(); } while (conv_hi < h1_0);
  }
  __syncthreads();
  if constexpr (TMA) { if (tid == 0) tma_refill(); }

  // w-interpolated planes (pairs along k) of two input rows per column; X / Y swap roles from one output row to the next, so
  // that the upper row of one output row is the lower row of the next without moving a register
  float2 wX[NC][V2_KW / 2], wY[NC][V2_KW / 2];
  int rowX = -1, rowY = -1;
  auto wrow = [&](int row, float2 (&w)[NC][V2_KW / 2]) {   // planes beyond the run read slack / stale words: never used
    const uint32_t so = (uint32_t)(row & (V2_SLOTS - 1)) * slot_bytes;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int q = 0; q < V2_KW / 4; ++q) {
        const uint4 a = lds128u(woff_a[c] + so + 16 * q), b = lds128u(woff_b[c] + so + 16 * q);
        w[c][2 * q] = ptx::ffma2(make_float2(__uint_as_float(a.x), __uint_as_float(a.y)), wa[c],
                                 ptx::fmul2(make_float2(__uint_as_float(b.x), __uint_as_float(b.y)), wb[c]));
        w[c][2 * q + 1] = ptx::ffma2(make_float2(__uint_as_float(a.z), __uint_as_float(a.w)), wa[c],
                                     ptx::fmul2(make_float2(__uint_as_float(b.z), __uint_as_float(b.w)), wb[c]));
      }
    }
  };
  const int oplane = p.tH * p.tW;                 // < 2^31 / 6 (host-checked): 32-bit element offsets
  float* optr = p.out + (long long)blockIdx.z * p.obatch + (long long)(od_base + zbeg - p.wd0 + p.pd0) * oplane +
                (long long)(oh_base - p.wh0 + p.ph0) * p.tW + (ow_base + lane - p.ww0 + p.pw0);
  float4 ht = s_htap[0];
  // one output row: lo / hi are the register planes that hold (or receive) rows h0 / h1
  auto out_row = [&](int t, float2 (&lo)[NC][V2_KW / 2], int& row_lo, float2 (&hi)[NC][V2_KW / 2], int& row_hi) {
    const int h0 = __float_as_int(ht.x), h1 = __float_as_int(ht.w);
    const float2 vh0 = make_float2(ht.y, ht.y), vh1 = make_float2(ht.z, ht.z);
    ht = s_htap[t + 1];                           // next row's taps: the load is in flight during this row's arithmetic
    if (active) {
      if constexpr (!kTapsInRegs) {
#pragma unroll
        for (int i = 0; i < V2_OPT; ++i) dt[i] = *reinterpret_cast<const float2*>(dtp + i);   // entries past the run: unused
      }
      if (row_lo != h0) { wrow(h0, lo); row_lo = h0; }
      if (row_hi != h1) {
        if (h1 != h0) {
          wrow(h1, hi);
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int q = 0; q < V2_KW / 2; ++q) hi[c][q] = lo[c][q];
        }
        row_hi = h1;
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float P[V2_KW];
#pragma unroll
        for (int q = 0; q < V2_KW / 2; ++q) {
          const float2 v = ptx::ffma2(lo[c][q], vh0, ptx::fmul2(hi[c][q], vh1));
          P[2 * q] = v.x; P[2 * q + 1] = v.y;
        }
        if (c == 0 || ow_ok[c]) {
          float* oc = optr + c * fw;
          if (quad43) {
#pragma unroll
            for (int i = 0; i < V2_OPT; ++i) {
              const int kk = (i / 3) * 4 + (i % 3);
              oc[i * oplane] = combine(P[kk], dt[i].x, P[kk + 1], dt[i].y);
            }
          } else {
            int i = 0;
#pragma unroll
            for (int kk = 0; kk < V2_KW - 1; ++kk) {
              if (emask & (1u << kk)) {           // warp-uniform; the i-th output of the run
                float2 tp = dt[0];
#pragma unroll
                for (int s2 = 1; s2 < V2_OPT; ++s2)
                  if (i == s2) tp = dt[s2];
                oc[i * oplane] = combine(P[kk], tp.x, P[kk + 1], tp.y);
                ++i;
              }
            }
          }
        }
      }
      optr += p.tW;
    }
    // rows of the next output row that are not converted yet (CTA-uniform): store, then ONE barrier
    const int g1 = __float_as_int(ht.w);
    if (g1 > conv_hi) {
      do { convert_next(); } while (conv_hi < g1);
      __syncthreads();
      if constexpr (TMA) { if (tid == 0) tma_refill(); }   // the slots just read are free: fetch the rows 4 ahead
    }
  };
  int t = 0;
  for (; t + 1 < n_oh; t += 2) {
    out_row(t, wX, rowX, wY, rowY);
    out_row(t + 1, wY, rowY, wX, rowX);
  }
  if (t < n_oh) out_row(t, wX, rowX, wY, rowY);
  if constexpr (!TMA) cp_async_wait<0>();   // nothing of this CTA is in flight when it exits (TMA: every fetched row was awaited)
}

__global__ void __launch_bounds__(256)
fill_f32_kernel(float4* __restrict__ x, long long nvec, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x)
    x[i] = make_float4(v, v, v, v);
}

// host mirror of taps() for sizing the shared tile (same fp32 operations)
void taps_host(int in, int out, int o, int& i0, int& i1) {
  if (in == out) { i0 = i1 = o; return; }
  const float scale = (float)in / (float)out;
  float src = fmaf(scale, (float)o + 0.5f, -0.5f);
  if (src < 0.f) src = 0.f;
  int a = (int)floorf(src);
  if (a > in - 1) a = in - 1;
  i0 = a;
  i1 = a + (a < in - 1 ? 1 : 0);
}
int max_span(int in, int out, int lo, int n, int tile) {
  int best = 1;
  for (int b = lo; b < lo + n; b += tile) {
    const int e = (b + tile < lo + n ? b + tile : lo + n) - 1;
    int a0, a1, b0, b1;
    taps_host(in, out, b, a0, a1);
    taps_host(in, out, e, b0, b1);
    if (b1 - a0 + 1 > best) best = b1 - a0 + 1;
  }
  return best;
}

}  // namespace

static int launch_fast(PrepParams p, int batch, float* lut_ws, int lut_lo, int lut_n, cudaStream_t s);

extern "C" int ctclip_prep_resample(const ctclip_prep_desc* d, void* stream) {
  if (d == nullptr) return ctclip::fail(CTCLIP_E_SHAPE, "prep_resample: null descriptor");
  if (d->batch <= 0) return CTCLIP_OK;
  if (d->D <= 0 || d->H <= 0 || d->W <= 0 || d->oD <= 0 || d->oH <= 0 || d->oW <= 0)
    return ctclip::fail(CTCLIP_E_SHAPE, "prep_resample: empty volume");
  int rc = ctclip::require_sm100();
  if (rc) return rc;
  PrepParams p{};
  p.in = d->in; p.out = d->out; p.in_is_i16 = d->in_is_i16; p.slope = d->slope; p.intercept = d->intercept;
  p.D = d->D; p.H = d->H; p.W = d->W;
  p.sd = d->stride_d; p.sh = d->stride_h; p.sw = d->stride_w; p.sbatch = d->stride_batch;
  p.oD = d->oD; p.oH = d->oH; p.oW = d->oW;
  if (d->in_is_i16 && (d->pre_op || d->post_op))
    return ctclip::fail(CTCLIP_E_SHAPE, "prep_resample: pre_op / post_op apply to fp32 input only");
  if ((d->pre_op != 0 && d->pre_op != 2 && d->pre_op != 3) || (d->post_op != 0 && d->post_op != 1))
    return ctclip::fail(CTCLIP_E_SHAPE, "prep_resample: unknown pre_op %d / post_op %d", d->pre_op, d->post_op);
  p.pre_op = d->pre_op; p.post_op = d->post_op; p.slope32 = (float)d->slope; p.icpt32 = (float)d->intercept;
  p.tD = d->tD > 0 ? d->tD : d->oD; p.tH = d->tH > 0 ? d->tH : d->oH; p.tW = d->tW > 0 ? d->tW : d->oW;
  // centre crop / pad split exactly as data.py:155-189 (python floor division)
  const int o[3] = {p.oD, p.oH, p.oW}, t[3] = {p.tD, p.tH, p.tW};
  int start[3], len[3], before[3];
  bool padded = false;
  for (int a = 0; a < 3; ++a) {
    int diff = o[a] - t[a];
    int s = diff >= 0 ? diff / 2 : -((-diff + 1) / 2);  // floor(diff / 2)
    int st = s > 0 ? s : 0;
    int en = s + t[a] < o[a] ? s + t[a] : o[a];
    start[a] = st; len[a] = en - st; before[a] = (t[a] - len[a]) / 2;
    if (len[a] != t[a]) padded = true;
  }
  p.wd0 = start[0]; p.wh0 = start[1]; p.ww0 = start[2];
  p.wdn = len[0]; p.whn = len[1]; p.wwn = len[2];
  p.pd0 = before[0]; p.ph0 = before[1]; p.pw0 = before[2];
  p.obatch = (long long)p.tD * p.tH * p.tW;
  p.kn_max = max_span(p.D, p.oD, p.wd0, p.wdn, TD);
  p.jn_max = max_span(p.W, p.oW, p.ww0, p.wwn, TW);
  // HU normalisation table (int16 input): index range where slope*v+intercept is not clipped, +-1 guard entries
  int lut_lo = 0, lut_n = 0;
  if (p.in_is_i16 && d->lut_workspace != nullptr && p.slope != 0.0) {
    const double a = (-1000.0 - p.intercept) / p.slope, b = (1000.0 - p.intercept) / p.slope;
    double lo = floor(a < b ? a : b) - 2.0, hi = ceil(a < b ? b : a) + 2.0;
    if (lo < -32768.0) lo = -32768.0;
    if (hi > 32767.0) hi = 32767.0;
    if (hi >= lo && hi - lo + 1.0 <= 8192.0) {
      lut_lo = (int)lo;
      lut_n = (int)(hi - lo + 1.0);
    }
  }
  const size_t smem = ((size_t)3 * p.kn_max * (p.jn_max | 1) + (size_t)lut_n) * sizeof(float);
  if (smem > 200 * 1024) return ctclip::fail(CTCLIP_E_SHAPE, "prep_resample: down-sampling factor too large for the shared tile (%zu B)", smem);
  cudaStream_t s = (cudaStream_t)stream;
  if (padded) {
    const long long n = p.obatch * d->batch;
    if (n % 4) return ctclip::fail(CTCLIP_E_ALIGN, "prep_resample: padded destination must have a multiple of 4 elements");
    fill_f32_kernel<<<ctclip::sm_count() * 8, 256, 0, s>>>((float4*)p.out, n / 4, d->pad_value);
    rc = ctclip::check_launch("prep_fill");
    if (rc) return rc;
  }
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(prep_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "prep_resample: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  if (lut_n > 0) {
    prep_lut_kernel<<<(lut_n + 255) / 256, 256, 0, s>>>(d->lut_workspace, lut_lo, lut_n, p.slope, p.intercept);
    rc = ctclip::check_launch("prep_lut");
    if (rc) return rc;
  }
  if (lut_n > 0 && !d->force_generic) {
    rc = launch_fast(p, d->batch, d->lut_workspace, lut_lo, lut_n, s);
    if (rc <= 0) return rc;
  }
  dim3 grid((unsigned)((p.whn + OHB - 1) / OHB), (unsigned)(((p.wdn + TD - 1) / TD) * ((p.wwn + TW - 1) / TW)),
            (unsigned)d->batch);
  prep_resample_kernel<<<grid, kThreads, smem, s>>>(p, d->lut_workspace, lut_lo, lut_n);
  return ctclip::check_launch("prep_resample");
}

// fast path (see prep_hwn_i16_kernel); returns 1 when it does not apply
template <bool ARITH, int NC>
static int launch_fast_t(const PrepParams& p, dim3 grid, size_t smem, float* lut_ws, int lut_lo, int lut_n, int icpt,
                         int ftd, int fw, int kn8, cudaStream_t s) {
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(prep_hwn_i16_kernel<ARITH, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return 1;
    configured = smem;
  }
  prep_hwn_i16_kernel<ARITH, NC><<<grid, 256, smem, s>>>(p, lut_ws, lut_lo, lut_n, icpt, ftd, fw, kn8);
  return ctclip::check_launch("prep_resample(hwn/i16)");
}

template <int NC>
static int launch_fast_nc(PrepParams p, int batch, float* lut_ws, int lut_lo, int lut_n, cudaStream_t s);

// v2 (prep_hwn_i16_v2_kernel): slope 1, integer intercept, every warp run of 6 output depths inside an aligned window of 8
// input planes, the CTA's input columns within 32; returns 1 when it does not apply. CTCLIP_PREP_V2=0 disables it (A/B).
static int launch_fast_v2(PrepParams p, int batch, cudaStream_t s) {
  const char* e = getenv("CTCLIP_PREP_V2");
  if (e != nullptr && e[0] == '0') return 1;
  if (!p.in_is_i16 || p.sd != 1 || p.sw != p.D || (p.D % 8) || (p.sh % 8) || (p.sbatch % 8) ||
      (reinterpret_cast<uintptr_t>(p.in) % 16) || p.D < p.oD)
    return 1;
  if (!(p.slope == 1.0 && p.intercept == floor(p.intercept) && fabs(p.intercept) <= 30000.0)) return 1;
  constexpr int ftd = FG * V2_OPT;
  // every warp run [b, b + 6): planes k0(first) & ~3 ... must cover k0(z) + 1 of every output of the run, each plane the
  // lower tap of at most one output (depth down-sampling), k0(z) within the first 7 planes of the window
  for (int b0 = p.wd0; b0 < p.wd0 + p.wdn; b0 += ftd)
    for (int g = 0; g < FG; ++g) {
      const int zb = b0 + g * V2_OPT;
      const int ze = (zb + V2_OPT < b0 + ftd ? zb + V2_OPT : b0 + ftd) < p.wd0 + p.wdn ? (zb + V2_OPT < b0 + ftd ? zb + V2_OPT : b0 + ftd)
                                                                                       : p.wd0 + p.wdn;
      if (zb >= ze) continue;
      int a0, a1, prev = -1;
      taps_host(p.D, p.oD, zb, a0, a1);
      const int w0 = a0 & ~3;
      for (int z = zb; z < ze; ++z) {
        taps_host(p.D, p.oD, z, a0, a1);
        if (a0 <= prev || a0 + 1 > w0 + V2_KW - 1) return 1;
        prev = a0;
      }
    }
  int kn8 = 8;
  for (int b = p.wd0; b < p.wd0 + p.wdn; b += ftd) {
    const int e2 = (b + ftd < p.wd0 + p.wdn ? b + ftd : p.wd0 + p.wdn) - 1;
    int a0, a1, b0, b1;
    taps_host(p.D, p.oD, b, a0, a1);
    taps_host(p.D, p.oD, e2, b0, b1);
    int hi = b1 > b0 + 1 ? b1 : b0 + 1;
    if (hi > p.D - 1) hi = p.D - 1;
    const int n8 = ((hi - (a0 & ~7)) / 8 + 1) * 8;
    if (n8 > kn8) kn8 = n8;
  }
  int fw = 32;
  while (fw >= 16 && max_span(p.W, p.oW, p.ww0, p.wwn, fw) > 32) --fw;
  if (fw < 16) return 1;
  const char* ne = getenv("CTCLIP_PREP_V2_NC");    // output columns per lane: 1 (default) | 2 (A/B: at the 128-register cap it spills)
  int nc = (ne != nullptr && ne[0] == '2') ? 2 : 1;
  auto chunk_slots = [&](int cols) { return (kn8 / 8) * ((cols + 1) & ~1); };   // columns are dealt to the threads in pairs
  if (nc == 2 && chunk_slots(max_span(p.W, p.oW, p.ww0, p.wwn, 2 * fw)) > V2_CPR * 256) nc = 1;
  const int jn = max_span(p.W, p.oW, p.ww0, p.wwn, nc * fw);
  if (chunk_slots(jn) > V2_CPR * 256) return 1;
  const int KP = kn8 + 4;
  // a warp's window may start up to 3 planes before its first plane and always spans 8: stays inside [0, KP + slack)
  // the row pipeline consumes input rows in increasing order without gaps: every row between the first and the last of a
  // brick is a tap of some output row, which holds whenever the height is not down-sampled by more than 2
  if ((long long)p.H > 2LL * p.oH) return 1;
  const int ring_chunks = ((chunk_slots(jn) + 255) / 256) * 256;   // 16-byte chunks of one raw row, whole 256-thread passes
  // raw rows by cp.async (default) or, CTCLIP_PREP_V2_TMA=1, by one TMA box per row when the brick fits it. Measured on the
  // production shape (32 scans): cp.async 3.05-3.25 ms, TMA 3.45 ms — with the LSU no longer the limiter (dependency
  // latency at 16 warps per SM is), the single refilling thread behind the barrier costs more than the LSU wavefronts it saves
  const char* te = getenv("CTCLIP_PREP_V2_TMA");
  const bool tma = (te != nullptr && te[0] == '1') && kn8 <= 64 && jn <= 32 * nc && ring_chunks * 16 == 128 * 32 * nc;
  const size_t smem = ((size_t)V2_SLOTS * jn * KP + 16) * sizeof(float) + (size_t)(ftd + FOH + 1) * 16 +
                      (size_t)V2_RING * ring_chunks * 16 + (tma ? 1024 + 64 : 0);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (tma) {   // raw scans as a 4-D tensor (depth, width, height, batch) of 16-bit words; box = 64 planes x 32 * nc columns of ONE row
    const cuuint64_t dims[4] = {(cuuint64_t)p.D, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {(cuuint64_t)p.sw * 2, (cuuint64_t)p.sh * 2, (cuuint64_t)p.sbatch * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)(32 * nc), 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (ctclip::encode_tmap(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(p.in), dims, strides, box, estr,
                            CU_TENSOR_MAP_SWIZZLE_128B) != CTCLIP_OK)
      return 1;
  }
  if (smem > 112 * 1024) return 1;
  if ((long long)p.tH * p.tW * V2_OPT >= (1LL << 31)) return 1;
  dim3 grid((unsigned)((p.whn + FOH - 1) / FOH), (unsigned)(((p.wdn + ftd - 1) / ftd) * ((p.wwn + nc * fw - 1) / (nc * fw))),
            (unsigned)batch);
  const char* oe = getenv("CTCLIP_PREP_V2_OCC");   // resident CTAs per SM the kernel is compiled for (register cap): 2 | 3
  const int occ = (oe != nullptr && oe[0] == '3') ? 3 : 2;
  auto launch = [&](auto kern, size_t& configured) -> int {
    if (smem > configured) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
      configured = smem;
    }
    kern<<<grid, 256, smem, s>>>(tmap, p, (int)p.intercept, fw, kn8, jn, ring_chunks);
    return ctclip::check_launch("prep_resample(hwn/i16 v2)");
  };
  static size_t conf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (nc == 2) return tma ? launch(prep_hwn_i16_v2_kernel<2, 2, true, 2>, conf[0]) : launch(prep_hwn_i16_v2_kernel<2, 2, false, 2>, conf[1]);
  if (ring_chunks > 256) return launch(prep_hwn_i16_v2_kernel<1, 2, false, 2>, conf[2]);   // two chunks per thread and row
  if (tma) return occ == 3 ? launch(prep_hwn_i16_v2_kernel<1, 3, true, 1>, conf[3]) : launch(prep_hwn_i16_v2_kernel<1, 2, true, 1>, conf[4]);
  return occ == 3 ? launch(prep_hwn_i16_v2_kernel<1, 3, false, 1>, conf[5]) : launch(prep_hwn_i16_v2_kernel<1, 2, false, 1>, conf[6]);
}

static int launch_fast(PrepParams p, int batch, float* lut_ws, int lut_lo, int lut_n, cudaStream_t s) {
  {
    const int rc = launch_fast_v2(p, batch, s);
    if (rc <= 0) return rc;
  }
  const char* e = getenv("CTCLIP_PREP_X2");   // 0: one output column per lane only (A/B and test hook)
  if (!(e != nullptr && e[0] == '0')) {
    const int rc = launch_fast_nc<2>(p, batch, lut_ws, lut_lo, lut_n, s);
    if (rc <= 0) return rc;
  }
  return launch_fast_nc<1>(p, batch, lut_ws, lut_lo, lut_n, s);
}

template <int NC>
static int launch_fast_nc(PrepParams p, int batch, float* lut_ws, int lut_lo, int lut_n, cudaStream_t s) {
  constexpr int FKMAX = FastCfg<NC>::KMAX, FCPR = FastCfg<NC>::CPR, FJP = FastCfg<NC>::JP;
  if (!p.in_is_i16 || lut_n <= 0 || p.sd != 1 || p.sw != p.D || (p.D % 8) || (p.sh % 8) || (p.sbatch % 8) ||
      (reinterpret_cast<uintptr_t>(p.in) % 16) || p.D < p.oD /* the k-indexed tap table needs depth down-sampling */)
    return 1;
  // depth run per warp: as long as the input planes it spans fit the FKMAX register planes
  int opt = NC == 1 ? 12 : 7;
  auto planes_ok = [&](int o) {
    for (int b = p.wd0; b < p.wd0 + p.wdn; b += o) {
      const int e = (b + o < p.wd0 + p.wdn ? b + o : p.wd0 + p.wdn) - 1;
      int a0, a1, b0, b1;
      taps_host(p.D, p.oD, b, a0, a1);
      taps_host(p.D, p.oD, e, b0, b1);
      int hi = b0 + 1 < p.D - 1 ? b0 + 1 : p.D - 1;
      if (hi - a0 + 1 > FKMAX) return false;
    }
    return true;
  };
  while (opt > 1 && !planes_ok(opt)) --opt;
  if (opt < 4) return 1;  // strong depth down-sampling: generic kernel
  // prefer a run length that tiles the kept depth range without a ragged last brick
  for (int o = opt; o >= opt - 3 && o >= 4; --o)
    if (p.wdn % (o * FG) == 0) { opt = o; break; }
  const int ftd = opt * FG;
  // shared tile extents over all bricks
  int kn8 = 8;
  for (int b = p.wd0; b < p.wd0 + p.wdn; b += ftd) {
    const int e = (b + ftd < p.wd0 + p.wdn ? b + ftd : p.wd0 + p.wdn) - 1;
    int a0, a1, b0, b1;
    taps_host(p.D, p.oD, b, a0, a1);
    taps_host(p.D, p.oD, e, b0, b1);
    int hi = b1 > b0 + 1 ? b1 : b0 + 1;
    if (hi > p.D - 1) hi = p.D - 1;
    const int n8 = ((hi - (a0 & ~7)) / 8 + 1) * 8;
    if (n8 > kn8) kn8 = n8;
  }
  // widest warp row (<= 32 outputs) whose input columns fit the 32 shared-memory banks
  int fw = 32;
  while (fw >= 16 && max_span(p.W, p.oW, p.ww0, p.wwn, fw) > 32) --fw;
  if (fw < 16) return 1;
  const int jn = max_span(p.W, p.oW, p.ww0, p.wwn, NC * fw);
  if ((kn8 / 8) * jn > FCPR * 256 || jn > FJP - 1) return 1;
  p.jn_max = jn;
  const bool arith = p.slope == 1.0 && p.intercept == floor(p.intercept) && fabs(p.intercept) <= 32768.0;
  const size_t smem = (((size_t)3 * (kn8 + 1) * FJP + FKMAX * FJP + 3) & ~(size_t)3) * sizeof(float) +
                      (size_t)(ftd + FOH) * 16 + (arith ? 0 : (size_t)lut_n * sizeof(float));
  if (smem > 100 * 1024) return 1;
  dim3 grid((unsigned)((p.whn + FOH - 1) / FOH), (unsigned)(((p.wdn + ftd - 1) / ftd) * ((p.wwn + NC * fw - 1) / (NC * fw))),
            (unsigned)batch);
  const int icpt = arith ? (int)p.intercept : 0;
  return arith ? launch_fast_t<true, NC>(p, grid, smem, lut_ws, lut_lo, lut_n, icpt, ftd, fw, kn8, s)
               : launch_fast_t<false, NC>(p, grid, smem, lut_ws, lut_lo, lut_n, icpt, ftd, fw, kn8, s);
}
