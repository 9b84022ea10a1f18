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

/* ------------------------------------------------------------------------------------------------
 * Row-wise memory-bound kernels (one warp per token row, 128-bit accesses, shuffle reductions).
 * layernorm_fwd: attention.py:28-35 (gamma-only; beta=NULL), attention.py:47 / ctvit.py:173 (nn.LayerNorm, eps 1e-5).
 *   y = (x-mean)*rstd*gamma (+beta); any of y_bf16 / raw_bf16 (bf16 copy of x) / y_f32 may be NULL.
 * layernorm_bwd: dx_out = (add_in?) + dLN(dy); optional bf16 copy; dgamma/dbeta accumulated with fp32 atomics.
 * geglu: attention.py:39-42 on a [rows][2*ld_half] bf16 buffer laid out [x | gate]. */
int ctclip_layernorm_fwd(const float* x, long long rows, int dim, const float* gamma, const float* beta, float eps,
                         void* y_bf16, void* raw_bf16, float* y_f32, void* stream);
int ctclip_layernorm_bwd(const float* dy, const float* x, long long rows, int dim, const float* gamma, float eps,
                         const float* add_in, float* dx_out, void* dx_bf16, float* dgamma, float* dbeta, void* stream);
int ctclip_geglu_fwd(const void* h, void* u, long long rows, int ld_half, void* stream);
int ctclip_geglu_bwd(const void* h, const void* du, void* dh, long long rows, int ld_half, void* stream);
int ctclip_cast_f32_bf16(const float* x, void* y, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PEG (attention.py:56-84): y = x + dwconv3x3x3_causal(x) + bias on tokens kept in the canonical (b,t,h,w,d) layout.
 * w27 is the depth-wise weight re-laid as [27][dim] (tap = (kt*3+kh)*3+kw). temporal=1 reproduces the reference's
 * reshape of the '(b h w) t d' tensor to (b,t,h,w,d) (ctvit.py:325-327) by index arithmetic. */
int ctclip_peg_fwd(const float* x, float* y, const float* w27, const float* bias, int batch, int t, int h, int w,
                   int dim, int temporal, void* stream);
int ctclip_peg_bwd_data(const float* dy, float* dx, void* dx_bf16, const float* w27, int batch, int t, int h, int w,
                        int dim, int temporal, void* stream);
int ctclip_peg_bwd_weight(const float* x, const float* dy, float* dw27, float* dbias, int batch, int t, int h, int w,
                          int dim, int temporal, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cosine-sim attention (attention.py:127-181), dim_head 32, tcgen05/TMEM. Tokens in canonical (b,t,h,w) order.
 *   q  [tokens][ldq]  bf16 (heads*32 used columns), kv [tokens][ldkv] bf16 (k | v), o [tokens][ldo] bf16,
 *   lse [tokens][heads] fp32 (log2 domain). temporal=0: sequences are frames of h*w tokens, additive bias from
 *   bias_table [heads][(2h-1)(2w-1)] (entry (dy+h-1)*(2w-1)+(dx+w-1), d = query - key) with bias_rowmax [heads][h*w]
 *   = max over keys; temporal=1: sequences are the t tokens of one (b,h,w) column, no bias.
 * Backward: dq [tokens][ldq], dkv [tokens][ldkv] bf16 (gradients w.r.t. the un-normalised projections),
 *   dq_scale/dk_scale [32] and dbias_table [heads][(2h-1)(2w-1)] accumulated with fp32 atomics. */
typedef struct ctclip_attn_desc {
  int batch, t, h, w, heads, dim_head, temporal;
  const void* q; int ldq;
  const void* kv; int ldkv;
  void* o; int ldo;
  float* lse;
  const float* q_scale; const float* k_scale;
  const float* bias_table; const float* bias_rowmax;
  /* backward only */
  const void* d_o;
  void* dq; void* dkv;
  float* dq_scale; float* dk_scale; float* dbias_table;
} ctclip_attn_desc;
int ctclip_attn_fwd(const ctclip_attn_desc* d, void* stream);
int ctclip_attn_bwd(const ctclip_attn_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCLIP_B200_H_ */
