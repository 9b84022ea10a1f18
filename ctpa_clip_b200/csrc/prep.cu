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
            v = reinterpret_cast<const float*>(p.in)[off + k];
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
            v = reinterpret_cast<const float*>(p.in)[off + (long long)j * p.sw];
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
      const float v = combine(ab, s_wd0[z], cd, s_wd1[z]);
      const int dd = od_base + z - p.wd0 + p.pd0;
      const int dw = ow_base + x - p.ww0 + p.pw0;
      out_b[((long long)dd * p.tH + dh) * p.tW + dw] = v;
    }
    // the next output row may overwrite the slot of a row that is no longer needed: wait for all readers
    __syncthreads();
  }
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
  dim3 grid((unsigned)((p.whn + OHB - 1) / OHB), (unsigned)(((p.wdn + TD - 1) / TD) * ((p.wwn + TW - 1) / TW)),
            (unsigned)d->batch);
  prep_resample_kernel<<<grid, kThreads, smem, s>>>(p, d->lut_workspace, lut_lo, lut_n);
  return ctclip::check_launch("prep_resample");
}
