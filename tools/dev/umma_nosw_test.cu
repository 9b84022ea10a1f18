// dev probe: tcgen05.mma with SWIZZLE_NONE core-matrix operands written by threads (no TMA),
// all four major-ness combinations. Standalone executable; prints max abs error vs a host reference.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../ctpa_clip_b200/csrc/ptx.cuh"
using namespace ptx;

__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// X[rows][cols] stored as 8x8 core matrices: addr(r,c) = (r/8)*RS + (c/8)*128 + (r%8)*16 + (c%8)*2
template <int M, int N, int K, bool A_MN, bool B_MN>
__global__ void test_kernel(const float* A /*M x K*/, const float* B /*N x K*/, float* C /*M x N*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // A stored as matrix [M rows][K cols] if !A_MN else stored as [K rows][M cols]
  constexpr int A_R = A_MN ? K : M, A_C = A_MN ? M : K;
  constexpr int B_R = B_MN ? K : N, B_C = B_MN ? N : K;
  constexpr int A_RS = (A_C / 8) * 128, B_RS = (B_C / 8) * 128;
  uint8_t* sa = smem;
  uint8_t* sb = smem + A_R * A_C * 2;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < A_R * A_C; i += blockDim.x) {
    int r = i / A_C, c = i % A_C;
    float v = A_MN ? A[c * K + r] : A[r * K + c];
    *reinterpret_cast<__nv_bfloat16*>(sa + (r / 8) * A_RS + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2) = __float2bfloat16_rn(v);
  }
  for (int i = threadIdx.x; i < B_R * B_C; i += blockDim.x) {
    int r = i / B_C, c = i % B_C;
    float v = B_MN ? B[c * K + r] : B[r * K + c];
    *reinterpret_cast<__nv_bfloat16*>(sb + (r / 8) * B_RS + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2) = __float2bfloat16_rn(v);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(M, N, A_MN, B_MN);
    for (int k = 0; k < K / 16; ++k) {
      // K-major: 2 core matrices along K are 128 B apart (LBO), 8-row groups RS apart (SBO); k-step = +256 B
      // MN-major: stored [K rows][MN cols]: mn groups 128 B apart (SBO), k groups (8 rows) RS apart (LBO); k-step = +2*RS
      uint64_t da = A_MN ? make_desc_nosw(smem_u32(sa) + k * 2 * A_RS, A_RS, 128) : make_desc_nosw(smem_u32(sa) + k * 256, 128, A_RS);
      uint64_t db = B_MN ? make_desc_nosw(smem_u32(sb) + k * 2 * B_RS, B_RS, 128) : make_desc_nosw(smem_u32(sb) + k * 256, 128, B_RS);
      mma_f16_ss(tmem, da, db, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp < 4) {
    int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 16) {
      uint32_t r[16];
      tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
      tmem_wait_ld();
      if (row < M) for (int j = 0; j < 16; ++j) C[row * N + c + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

static float bf(float x) { __nv_bfloat16 b = __float2bfloat16_rn(x); return __bfloat162float(b); }

template <int M, int N, int K, bool A_MN, bool B_MN>
int run() {
  std::vector<float> A(M * K), B(N * K), C(M * N), R(M * N);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.f;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)bf(A[m * K + k]) * bf(B[n * K + k]); R[m * N + n] = (float)s; }
  float *dA, *dB, *dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dC, 0, C.size() * 4);
  int smem = (M * K + N * K) * 2;
  cudaFuncSetAttribute(test_kernel<M, N, K, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  test_kernel<M, N, K, A_MN, B_MN><<<1, 128, smem>>>(dA, dB, dC);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0; for (int i = 0; i < M * N; ++i) err = fmax(err, fabs(C[i] - R[i]));
  printf("M=%d N=%d K=%d A_MN=%d B_MN=%d : %s maxerr=%g\n", M, N, K, (int)A_MN, (int)B_MN, cudaGetErrorString(e), err);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return (e == cudaSuccess && err < 1e-3) ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<128, 128, 32, false, false>();   // S = Q K^T
  bad += run<128, 32, 128, false, true>();    // O = P V        (V [keys][d] as MN-major B)
  bad += run<128, 32, 128, true, true>();     // dV = P^T dO    (both MN-major)
  bad += run<128, 128, 32, true, false>();
  bad += run<128, 64, 64, false, false>();
  bad += run<128, 48, 48, false, true>();     // N not multiple of 32
  printf(bad ? "SOME_FAILED\n" : "ALL_OK\n");
  return bad;
}
