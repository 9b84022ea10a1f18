/* TEST INFRASTRUCTURE — plain-C restatement of the data_prep volume normalisation (CPU oracle).
 *
 * Follows, line by line:
 *   hu_normalise   : CTPA_CLIP/data_prep/preprocess_train.py:99-104  (slope*x+intercept in float64, clip to
 *                    [-1000,1000], /1000, astype(float32), transpose(2,0,1))
 *   trilinear      : F.interpolate(mode='trilinear', align_corners=False) as called by resize_array,
 *                    CTPA_CLIP/data_prep/preprocess_train.py:31-42 == ct_clip/data.py:15-40. The index/weight rule
 *                    and the per-level combine  r = fma(t0, w0, t1*w1)  were pinned bit-for-bit against torch 2.11
 *                    CPU (tools/make_golden.py writes tests/golden/resample_*.npz from the reference's resize_array).
 *   crop_pad       : CTPA_CLIP/ct_clip/data.py:155-190 (centre crop to <=(480,480,240), pad with -1, permute)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call this.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/build_oracle.py); -ffp-contract=off so that the only
 * fused operations are the explicit fmaf() calls below.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* raw is (H, W, N) with N contiguous (NIfTI array order after get_fdata); out is (N, H, W) float32 */
void ctclip_oracle_hu_normalise_i16(const int16_t* raw, int H, int W, int N, double slope, double intercept,
                                    float* out) {
  for (int i = 0; i < H; ++i)
    for (int j = 0; j < W; ++j)
      for (int k = 0; k < N; ++k) {
        double v = slope * (double)raw[((size_t)i * W + j) * N + k];
        v = v + intercept;
        if (v < -1000.0) v = -1000.0;
        if (v > 1000.0) v = 1000.0;
        out[((size_t)k * H + i) * W + j] = (float)(v / 1000.0);
      }
}

void ctclip_oracle_taps(int in, int out, int o, int* i0, int* i1, float* w0, float* w1) {
  float scale = (float)in / (float)out;
  float src = fmaf(scale, (float)o + 0.5f, -0.5f); /* single rounding */
  if (src < 0.f) src = 0.f;
  int a = (int)floorf(src);
  if (a > in - 1) a = in - 1;
  float l = src - (float)a;
  if (l < 0.f) l = 0.f;
  if (l > 1.f) l = 1.f;
  *i0 = a;
  *i1 = a + (a < in - 1 ? 1 : 0);
  *w1 = l;
  *w0 = 1.f - l;
  if (in == out) { *i0 = o; *i1 = o; *w0 = 1.f; *w1 = 0.f; }
}

static inline float combine(float t0, float w0, float t1, float w1) { return fmaf(t0, w0, t1 * w1); }

/* in (D,H,W) float32 -> out (oD,oH,oW) float32 */
void ctclip_oracle_trilinear(const float* in, int D, int H, int W, float* out, int oD, int oH, int oW) {
  for (int d = 0; d < oD; ++d) {
    int d0, d1; float wd0, wd1;
    ctclip_oracle_taps(D, oD, d, &d0, &d1, &wd0, &wd1);
    for (int h = 0; h < oH; ++h) {
      int h0, h1; float wh0, wh1;
      ctclip_oracle_taps(H, oH, h, &h0, &h1, &wh0, &wh1);
      const float* p00 = in + ((size_t)d0 * H + h0) * W;
      const float* p01 = in + ((size_t)d0 * H + h1) * W;
      const float* p10 = in + ((size_t)d1 * H + h0) * W;
      const float* p11 = in + ((size_t)d1 * H + h1) * W;
      for (int w = 0; w < oW; ++w) {
        int w0, w1; float ww0, ww1;
        ctclip_oracle_taps(W, oW, w, &w0, &w1, &ww0, &ww1);
        float a = combine(p00[w0], ww0, p00[w1], ww1);
        float b = combine(p01[w0], ww0, p01[w1], ww1);
        float c = combine(p10[w0], ww0, p10[w1], ww1);
        float e = combine(p11[w0], ww0, p11[w1], ww1);
        float ab = combine(a, wh0, b, wh1);
        float ce = combine(c, wh0, e, wh1);
        out[((size_t)d * oH + h) * oW + w] = combine(ab, wd0, ce, wd1);
      }
    }
  }
}

/* in (D,H,W) float32 (already normalised) -> out (tD,tH,tW): centre crop then pad with pad_value, floor/ceil split.
 * data.py works on (H,W,D) and permutes to (D,H,W) at the end; the arithmetic per axis is identical. */
void ctclip_oracle_crop_pad(const float* in, int D, int H, int W, float* out, int tD, int tH, int tW,
                            float pad_value) {
  int sz[3] = {D, H, W}, tg[3] = {tD, tH, tW}, start[3], len[3], before[3];
  for (int a = 0; a < 3; ++a) {
    int s = (sz[a] - tg[a]) / 2; /* python floor-div then max(...,0) */
    if ((sz[a] - tg[a]) < 0 && ((sz[a] - tg[a]) % 2 != 0)) s -= 1;
    int st = s > 0 ? s : 0;
    int en = s + tg[a] < sz[a] ? s + tg[a] : sz[a];
    start[a] = st;
    len[a] = en - st;
    before[a] = (tg[a] - len[a]) / 2;
  }
  for (int d = 0; d < tD; ++d)
    for (int h = 0; h < tH; ++h)
      for (int w = 0; w < tW; ++w) {
        int sd = d - before[0], sh = h - before[1], sw = w - before[2];
        float v = pad_value;
        if (sd >= 0 && sd < len[0] && sh >= 0 && sh < len[1] && sw >= 0 && sw < len[2])
          v = in[((size_t)(start[0] + sd) * H + (start[1] + sh)) * W + (start[2] + sw)];
        out[((size_t)d * tH + h) * tW + w] = v;
      }
}

/* ---- DataLoader-side arithmetic (the two loaders that feed CT-CLIP) ------------------------------------------------
 * All float32, one rounding per numpy operation (NEP 50: the python scalars act as float32 against a float32 array).
 *   affine_f32   : CTPA_CLIP/ct_clip/data.py:138   img = slope * ct_scan + intercept           (before the resample)
 *   clip_div_f32 : CTPA_CLIP/ct_clip/data.py:150-152  clip(img, -1000, 1000) / 1000             (after the resample)
 *   window_infer : CTPA_CLIP/ct_clip/data_inference.py:81-85  clip(img * 1000, -1000, 200); (img + 400) / 600 */
void ctclip_oracle_affine_f32(const float* in, size_t n, float slope, float intercept, float* out) {
  for (size_t i = 0; i < n; ++i) {
    float v = slope * in[i];
    out[i] = v + intercept;
  }
}
void ctclip_oracle_clip_div_f32(const float* in, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) {
    float v = in[i];
    if (v < -1000.f) v = -1000.f;
    if (v > 1000.f) v = 1000.f;
    out[i] = v / 1000.f;
  }
}
void ctclip_oracle_window_infer(const float* in, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) {
    float v = in[i] * 1000.f;
    if (v < -1000.f) v = -1000.f;
    if (v > 200.f) v = 200.f;
    v = v + 400.f;
    out[i] = v / 600.f;
  }
}
