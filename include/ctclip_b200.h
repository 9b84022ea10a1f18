/* ctclip_b200.h — C-ABI of libctclip_sm100.so: the B200 (sm_100a) kernels behind the CT-CLIP hot path.
 *
 * The reference (sharonct/CTPA-CLIP) is pure Python/PyTorch and has no FFI of its own; every entry point
 * below replaces the stock ATen/cuBLAS/cuDNN call sequence behind one reference operator, cited per
 * function as CTPA_CLIP/<file>:<line>. The binding a maintainer adds on the reference side is a ctypes
 * stub; the in-tree host mirror lives in ctpa_clip_b200/.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types. All pointers are DEVICE pointers unless named host_*.
 *   - the library never allocates, frees or retains device memory; workspaces are caller-owned.
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous, no internal syncs.
 *   - return 0 on success, negative CTCLIP_E_* otherwise; text via ctclip_last_error (thread-local).
 *   - there is NO CPU fallback: on a non-sm_100 device every compute entry returns CTCLIP_E_ARCH.
 */
#ifndef CTCLIP_B200_H_
#define CTCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTCLIP_OK 0
#define CTCLIP_E_SHAPE (-1)
#define CTCLIP_E_ALIGN (-2)
#define CTCLIP_E_ARCH (-3)
#define CTCLIP_E_CUDA (-4)

/* library version (major*10000 + minor*100 + patch) */
int ctclip_version(void);
/* copies the calling thread's last error text into buf (NUL-terminated); returns its length */
int ctclip_last_error(char* buf, size_t n);
/* number of kernels launched by this process through the library since load (bench.py: gpu_launches) */
long long ctclip_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM  C[M,N] (op)= alpha * sum_k A(m,k) B(n,k) (+ bias[n]) (+ resid[m,n])      tcgen05 / TMEM / TMA
 * Replaces nn.Linear forward / dgrad / wgrad on the path: ctvit.py:172 (patch embed), attention.py:119,120
 * (to_q,to_kv), :125 (to_out), :48,:51 (FeedForward), ct_clip.py:549 (to_text_latent).
 *   a_mn_major = 0: A is [M][lda] with K contiguous;   1: A is [K][lda] with M contiguous (i.e. A^T stored)
 *   b_mn_major = 0: B is [N][ldb] with K contiguous;   1: B is [K][ldb] with N contiguous
 *   C is row-major [M][ldc], bf16 (c_is_f32=0) or fp32 (c_is_f32=1); resid is fp32 [M][ldr] (may alias C)
 *   atomic=1: C += result with fp32 atomics (required for splits>1); splits<=0 picks a split-K factor. */
typedef struct ctclip_gemm_desc {
  int M, N, K;
  const void* A; long long lda; int a_mn_major;
  const void* B; long long ldb; int b_mn_major;
  void* C; long long ldc; int c_is_f32;
  const float* bias;
  const float* resid; long long ldr;
  float alpha;
  int atomic;
  int splits;
} ctclip_gemm_desc;
int ctclip_gemm_bf16(const ctclip_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCLIP_B200_H_ */
