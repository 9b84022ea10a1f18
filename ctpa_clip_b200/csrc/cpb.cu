// Continuous position bias (reference attention.py:229-276): the 2 -> dim -> dim -> heads MLP with LeakyReLU(0.1) on the
// signed-log relative offsets of an h x w grid. The reference evaluates it on all (h*w)^2 query/key pairs (177 GF at
// 24 x 24); only R = (2h-1)(2w-1) offsets are distinct, so the table is evaluated on those R rows (1.2 GF) and the
// attention kernels gather from it. Forward + hand-written backward, fp32 throughout (the table feeds the logits of
// every spatial layer; parameter-sized work, so CUDA-core FFMA tiles: 64 x 64 x 16 shared-memory tiles, 4 x 4 per thread).
#include "ptx.cuh"
#include "ctclip_internal.h"

namespace {

constexpr float kSlope = 0.1f;  // leaky_relu(p=0.1), attention.py:20-21

// rel[r] = (dy, dx) with r = (dy+h-1)*(2w-1) + (dx+w-1); x = sign(v) * log(|v| + 1) (attention.py:266-267)
// h0[r][c] = lrelu(W0[c][0] x0 + W0[c][1] x1 + b0[c])
__global__ void __launch_bounds__(256)
cpb_layer0_kernel(int h, int w, int dim, int log_dist, const float* __restrict__ W0, const float* __restrict__ b0,
                  float* __restrict__ xin, float* __restrict__ h0) {
  const int R = (2 * h - 1) * (2 * w - 1);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)R * dim) return;
  const int r = (int)(idx / dim), c = (int)(idx % dim);
  float x0 = (float)(r / (2 * w - 1) - (h - 1));
  float x1 = (float)(r % (2 * w - 1) - (w - 1));
  if (log_dist) {
    x0 = (x0 > 0.f ? 1.f : (x0 < 0.f ? -1.f : 0.f)) * logf(fabsf(x0) + 1.f);
    x1 = (x1 > 0.f ? 1.f : (x1 < 0.f ? -1.f : 0.f)) * logf(fabsf(x1) + 1.f);
  }
  if (c == 0) {
    xin[2 * r] = x0;
    xin[2 * r + 1] = x1;
  }
  const float z = fmaf(W0[2 * c + 1], x1, fmaf(W0[2 * c], x0, b0[c]));
  h0[idx] = z > 0.f ? z : kSlope * z;
}

// C(m,n) = epilogue( sum_k A(m,k) B(k,n) ), all three operands through element strides.
//   bias (per n, may be NULL) is added first; act = 1 applies LeakyReLU; gate (same strides as C, may be NULL) multiplies
//   by LeakyReLU'(gate) = gate > 0 ? 1 : slope (the activation OUTPUT has the sign of its input).
struct SgemmArgs {
  int M, N, K;
  const float* A; long long a_sm, a_sk;
  const float* B; long long b_sk, b_sn;
  float* C; long long c_sm, c_sn;
  const float* bias;
  const float* gate;
  int act;
  int kchunk;  // k range per blockIdx.z (multiple of 16); gridDim.z > 1 -> partial sums go out as fp32 atomics into a zeroed C
};

__global__ void __launch_bounds__(256)
cpb_sgemm_kernel(SgemmArgs p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      // pick the faster-varying index so that the unit-stride operand dimension is the coalesced one
      int kk, mm;
      if (p.a_sk == 1) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < kend) ? p.A[m * p.a_sm + k * p.a_sk] : 0.f;
      int kb, nn;
      if (p.b_sk == 1) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
      const int n = n0 + nn, k2 = k0 + kb;
      Bs[kb][nn] = (n < p.N && k2 < kend) ? p.B[k2 * p.b_sk + n * p.b_sn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (gridDim.z > 1) {
        atomicAdd(p.C + m * p.c_sm + n * p.c_sn, v);
        continue;
      }
      if (p.bias != nullptr) v += p.bias[n];
      if (p.act) v = v > 0.f ? v : kSlope * v;
      if (p.gate != nullptr) v *= p.gate[m * p.c_sm + n * p.c_sn] > 0.f ? 1.f : kSlope;
      p.C[m * p.c_sm + n * p.c_sn] = v;
    }
  }
}

// out[n] += sum_m X(m, n) (bias gradients; out zeroed by the caller). Row-major X (sn == 1): a 32-column x 8-row-lane
// tile per CTA walks a chunk of rows with coalesced 128-byte reads; column-major X (sm == 1): one warp per column.
__global__ void __launch_bounds__(256)
cpb_colsum_kernel(const float* __restrict__ X, long long sm, long long sn, int M, int N, int rows_per_cta,
                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int mbeg = blockIdx.y * rows_per_cta, mend = min(M, mbeg + rows_per_cta);
  if (sn == 1) {
    __shared__ float red[8][33];
    const int n = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (n < N)
      for (int m = mbeg + wy; m < mend; m += 8) s += X[m * sm + n];
    red[wy][lane] = s;
    __syncthreads();
    if (wy == 0 && n < N) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][lane];
      atomicAdd(out + n, t);
    }
  } else {
    const int n = blockIdx.x * 8 + wy;
    if (n >= N) return;
    float s = 0.f;
    for (int m = mbeg + lane; m < mend; m += 32) s += X[m * sm + n * sn];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) atomicAdd(out + n, s);
  }
}

// table[head][r] = <h1[r], W2[head]> + b2[head], written transposed ('i j h -> h i j', attention.py:274): one warp per
// offset row r keeps its h1 row in registers and reduces it against every head (heads is tiny: a 64-wide GEMM tile
// would idle 7/8 of its columns).
template <int kMaxPerLane>
__global__ void __launch_bounds__(256)
cpb_out_kernel(const float* __restrict__ h1, const float* __restrict__ W2, const float* __restrict__ b2, int R, int dim,
               int heads, float* __restrict__ table) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  float x[kMaxPerLane];
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int k = lane + 32 * i;
    x[i] = k < dim ? h1[(long long)r * dim + k] : 0.f;
  }
  for (int hd = 0; hd < heads; ++hd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int k = lane + 32 * i;
      if (k < dim) s = fmaf(x[i], W2[(long long)hd * dim + k], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) table[(long long)hd * R + r] = s + b2[hd];
  }
}

// rowmax[head][q] = max over keys k of table[head][(qy-ky+h-1)*(2w-1) + (qx-kx+w-1)]: the running-max seed of the
// attention kernels' softmax. One warp per (head, q).
__global__ void __launch_bounds__(256)
cpb_rowmax_kernel(const float* __restrict__ table, int h, int w, int heads, float* __restrict__ rowmax) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n = h * w;
  if (item >= heads * n) return;
  const int head = item / n, q = item % n;
  const int qy = q / w, qx = q % w;
  const int R = (2 * h - 1) * (2 * w - 1);
  float m = -INFINITY;
  for (int k = lane; k < n; k += 32) {
    const int ky = k / w, kx = k % w;
    m = fmaxf(m, table[(long long)head * R + (qy - ky + h - 1) * (2 * w - 1) + (qx - kx + w - 1)]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) rowmax[item] = m;
}

int sgemm(SgemmArgs a, cudaStream_t s, const char* what, int splits = 1) {
  splits = max(1, min(splits, (a.K + 15) / 16));
  a.kchunk = (((a.K + splits - 1) / splits) + 15) / 16 * 16;
  splits = (a.K + a.kchunk - 1) / a.kchunk;
  if (splits > 1) {
    if (a.bias != nullptr || a.gate != nullptr || a.act || a.c_sn != 1 || a.c_sm != a.N)
      return ctclip::fail(CTCLIP_E_SHAPE, "%s: split-K needs a linear epilogue and a dense C", what);
    cudaError_t e = cudaMemsetAsync(a.C, 0, sizeof(float) * (size_t)a.M * a.N, s);
    if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "%s: memset: %s", what, cudaGetErrorString(e));
  }
  dim3 grid((a.N + 63) / 64, (a.M + 63) / 64, splits);
  cpb_sgemm_kernel<<<grid, 256, 0, s>>>(a);
  return ctclip::check_launch(what);
}

// out[N] = column sums of X (M rows)
int colsum(const float* X, long long sm, long long sn, int M, int N, float* out, cudaStream_t s, const char* what) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, s);
  if (e != cudaSuccess) return ctclip::fail(CTCLIP_E_CUDA, "%s: memset: %s", what, cudaGetErrorString(e));
  const int rows_per_cta = 128;
  dim3 grid(sn == 1 ? (N + 31) / 32 : (N + 7) / 8, (M + rows_per_cta - 1) / rows_per_cta);
  cpb_colsum_kernel<<<grid, 256, 0, s>>>(X, sm, sn, M, N, rows_per_cta, out);
  return ctclip::check_launch(what);
}

int check_dims(int h, int w, int dim, int heads, const char* who) {
  if (h <= 0 || w <= 0 || dim <= 0 || heads <= 0 || h > 1024 || w > 1024)
    return ctclip::fail(CTCLIP_E_SHAPE, "%s: bad shape h=%d w=%d dim=%d heads=%d", who, h, w, dim, heads);
  return ctclip::require_sm100();
}

}  // namespace

// acts: fp32 workspace [R*2 + 2*R*dim] = signed-log offsets x [R][2], h0 [R][dim], h1 [R][dim] (kept for the backward).
// table: fp32 [heads][R]; rowmax: fp32 [heads][h*w] or NULL.
extern "C" int ctclip_cpb_table_fwd(int h, int w, int dim, int heads, int log_dist, const float* W0, const float* b0,
                                    const float* W1, const float* b1, const float* W2, const float* b2, float* acts,
                                    float* table, float* rowmax, void* stream) {
  int rc = check_dims(h, w, dim, heads, "cpb_table_fwd");
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int R = (2 * h - 1) * (2 * w - 1);
  float* xin = acts;
  float* h0 = acts + 2 * (long long)R;
  float* h1 = h0 + (long long)R * dim;
  cpb_layer0_kernel<<<(unsigned)(((long long)R * dim + 255) / 256), 256, 0, s>>>(h, w, dim, log_dist, W0, b0, xin, h0);
  rc = ctclip::check_launch("cpb_layer0");
  if (rc) return rc;
  // h1 = lrelu(h0 W1^T + b1)
  SgemmArgs l1{R, dim, dim, h0, dim, 1, W1, 1, dim, h1, dim, 1, b1, nullptr, 1, 0};
  rc = sgemm(l1, s, "cpb_layer1");
  if (rc) return rc;
  // table[head][r] = h1 W2^T + b2, written transposed ('i j h -> h i j', attention.py:274)
  if (dim <= 512) {
    cpb_out_kernel<16><<<(R + 7) / 8, 256, 0, s>>>(h1, W2, b2, R, dim, heads, table);
    rc = ctclip::check_launch("cpb_layer2");
  } else {
    SgemmArgs l2{R, heads, dim, h1, dim, 1, W2, 1, dim, table, 1, R, b2, nullptr, 0, 0};
    rc = sgemm(l2, s, "cpb_layer2");
  }
  if (rc) return rc;
  if (rowmax != nullptr) {
    cpb_rowmax_kernel<<<(heads * h * w + 7) / 8, 256, 0, s>>>(table, h, w, heads, rowmax);
    rc = ctclip::check_launch("cpb_rowmax");
  }
  return rc;
}

// dtable: fp32 [heads][R] (summed over layers and sequences by the attention backward). acts: as written by the forward.
// work: fp32 [2*R*dim]. Outputs are OVERWRITTEN: dW0 [dim][2], db0 [dim], dW1 [dim][dim], db1 [dim], dW2 [heads][dim], db2 [heads].
extern "C" int ctclip_cpb_table_bwd(int h, int w, int dim, int heads, const float* W1, const float* W2, const float* acts,
                                    const float* dtable, float* work, float* dW0, float* db0, float* dW1, float* db1,
                                    float* dW2, float* db2, void* stream) {
  int rc = check_dims(h, w, dim, heads, "cpb_table_bwd");
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int R = (2 * h - 1) * (2 * w - 1);
  const float* xin = acts;
  const float* h0 = acts + 2 * (long long)R;
  const float* h1 = h0 + (long long)R * dim;
  float* dz1 = work;
  float* dz0 = work + (long long)R * dim;
  // dz2(r, head) = dtable[head][r]
  // dW2[head][k] = sum_r dz2(r, head) h1[r][k]
  const int ksplits = (R + 127) / 128;  // the three weight gradients reduce over the R offset rows: split-K fills the SMs
  SgemmArgs gw2{heads, dim, R, dtable, R, 1, h1, dim, 1, dW2, dim, 1, nullptr, nullptr, 0, 0};
  if ((rc = sgemm(gw2, s, "cpb_dW2", ksplits))) return rc;
  if ((rc = colsum(dtable, 1, R, R, heads, db2, s, "cpb_db2"))) return rc;
  // dz1[r][c] = (sum_head dz2(r, head) W2[head][c]) * lrelu'(h1[r][c])
  SgemmArgs g1{R, dim, heads, dtable, 1, R, W2, dim, 1, dz1, dim, 1, nullptr, h1, 0, 0};
  if ((rc = sgemm(g1, s, "cpb_dz1"))) return rc;
  // dW1[c][k] = sum_r dz1[r][c] h0[r][k]
  SgemmArgs gw1{dim, dim, R, dz1, 1, dim, h0, dim, 1, dW1, dim, 1, nullptr, nullptr, 0, 0};
  if ((rc = sgemm(gw1, s, "cpb_dW1", ksplits))) return rc;
  if ((rc = colsum(dz1, dim, 1, R, dim, db1, s, "cpb_db1"))) return rc;
  // dz0[r][k] = (sum_c dz1[r][c] W1[c][k]) * lrelu'(h0[r][k])
  SgemmArgs g0{R, dim, dim, dz1, dim, 1, W1, dim, 1, dz0, dim, 1, nullptr, h0, 0, 0};
  if ((rc = sgemm(g0, s, "cpb_dz0"))) return rc;
  // dW0[c][j] = sum_r dz0[r][c] x[r][j]
  SgemmArgs gw0{dim, 2, R, dz0, 1, dim, xin, 2, 1, dW0, 2, 1, nullptr, nullptr, 0, 0};
  if ((rc = sgemm(gw0, s, "cpb_dW0", ksplits))) return rc;
  return colsum(dz0, dim, 1, R, dim, db0, s, "cpb_db0");
}
