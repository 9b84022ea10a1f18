/* ctclip_b200.h — C-ABI of libctclip_sm100.so: the B200 (sm_100a) kernels behind the CT-CLIP hot path.
 *
 * The reference (sharonct/CTPA-CLIP) is pure Python/PyTorch and has no FFI of its own; every entry point
 * below replaces the stock ATen/cuBLAS/cuDNN call sequence behind one reference operator, cited per
 * function as CTPA_CLIP/<file>:<line>. The binding a maintainer adds on the reference side is a ctypes
 * stub; the in-tree host mirror lives in ctpa_clip_b200/.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types. All pointers are DEVICE pointers unless named host_*.
 *   - the library never allocates, frees or retains device memory; workspaces are caller-owned. The one
 *     exception are the explicitly registered symmetric buffers of the peer-memory latent exchange (ctclip_symm_*).
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
/* bytes of the caller-owned workspace of an entry point: op in {"clip_loss" (B, d), "clip_loss_allgather" (b_local, d, world),
 * "cpb_table_fwd" (h, w, dim) [the `acts` buffer], "cpb_table_bwd" (h, w, dim), "bert_attn_bwd" (batch, heads, seq_len),
 * "prep_resample" () [lut_workspace], "sumsq" ()}; -1 for an unknown op / wrong ndims. Pure host arithmetic. */
long long ctclip_workspace_bytes(const char* op, const long long* dims, int ndims);
/* number of kernels launched by this process through the library since load (bench.py: gpu_launches) */
long long ctclip_launch_count(void);
/* SMs the persistent kernels (GEMM, attention, ...) size their grids for: 0 = all (default; env CTCLIP_SM_BUDGET). Data-parallel
 * training leaves a few SMs to the NCCL kernels that overlap the backward pass: a static persistent schedule over every SM
 * would run a second wave for each SM a communication kernel holds. Process-wide; takes effect at the next launch. */
int ctclip_set_sm_budget(int sms);

/* ------------------------------------------------------------------------------------------------
 * GEMM  C[M,N] (op)= alpha * sum_k A(m,k) B(n,k) (+ bias[n]) (+ resid[m,n])      tcgen05 / TMEM / TMA
 * Replaces nn.Linear forward / dgrad / wgrad on the path: ctvit.py:172 (patch embed), attention.py:119,120
 * (to_q,to_kv), :125 (to_out), :48,:51 (FeedForward), ct_clip.py:549 (to_text_latent).
 *   a_mn_major = 0: A is [M][lda] with K contiguous;   1: A is [K][lda] with M contiguous (i.e. A^T stored)
 *   b_mn_major = 0: B is [N][ldb] with K contiguous;   1: B is [K][ldb] with N contiguous
 *   C is row-major [M][ldc], bf16 (c_is_f32=0) or fp32 (c_is_f32=1); resid is fp32 [M][ldr] (may alias C)
 *   atomic=1: C += result with fp32 atomics (required for splits>1); splits<=0 picks a split-K factor.
 *   top2_out: VQ-assignment epilogue (ctvit.py:427): instead of C, the two largest accumulators of every row within
 *   each N-tile and their column indices. */
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
  void* top2_out; /* non-NULL: float4 [M][ceil(N/BN)] = per n-tile best two (value, column-as-int-bits); C unused */
  /* batched mode (multi-head attention products of the BERT text tower, transformers BertSelfAttention): batch_h * batch_b
   * independent M x N x K problems; problem (b, h) reads A + b*a_stride_b + h*a_stride_h (elements; same for B, C).
   * M, N, K are per-problem extents: tiles never read a neighbouring problem (TMA zero fill). 0 / 1 = not batched. */
  int batch_h, batch_b;
  long long a_stride_h, a_stride_b, b_stride_h, b_stride_b, c_stride_h, c_stride_b;
  /* GEGLU epilogue of the FeedForward block's first Linear (attention.py:39-48); N = 2 * Nh, bf16 C, K-major A and B, no
   * bias / resid / split-K. geglu_u != NULL: B is [2 Nh][ldb] = [x rows | gate rows]; C [M][N] receives h = [x | gate] (kept
   * for the backward) and geglu_u [M][ld_u] receives x * gelu(gate), computed from the bf16-rounded h (bit-identical to
   * ctclip_geglu_fwd on C) without re-reading h from HBM.
   * geglu_h / ld_h: reserved for the backward counterpart (GEGLU backward in the FF2 dgrad epilogue). It was built and is
   * bit-identical, but the h loads in the epilogue make it slower than the two-kernel path (491 vs 192 + 250 us per layer),
   * so the library rejects a non-NULL geglu_h. */
  void* geglu_u; long long ld_u;
  const void* geglu_h; long long ld_h;
} ctclip_gemm_desc;
/* N-tile width the kernel will use for a given N (128 or 256): sizes the top2_out buffer */
int ctclip_gemm_tile_n(int N);
int ctclip_gemm_bf16(const ctclip_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise memory-bound kernels (one warp per token row, 128-bit accesses, shuffle reductions).
 * layernorm_fwd: attention.py:28-35 (gamma-only; beta=NULL), attention.py:47 / ctvit.py:173 (nn.LayerNorm, eps 1e-5).
 *   y = (x-mean)*rstd*gamma (+beta); any of y_bf16 / raw_bf16 (bf16 copy of x) / y_f32 may be NULL.
 * layernorm_bwd: dx_out = (add_in?) + dLN(dy); optional bf16 copy; dgamma/dbeta accumulated with fp32 atomics.
 * geglu: attention.py:39-42 on a [rows][2*ld_half] bf16 buffer laid out [x | gate]. */
int ctclip_layernorm_fwd(const float* x, long long rows, int dim, const float* gamma, const float* beta, float eps,
                         void* y_bf16, void* raw_bf16, float* y_f32, void* stream);
int ctclip_layernorm_bwd(const void* dy /* fp32, or bf16 when dy_is_bf16 (straight from a GEMM epilogue) */, int dy_is_bf16,
                         const float* x, long long rows, int dim, const float* gamma, float eps, const float* add_in,
                         float* dx_out, void* dx_bf16, float* dgamma, float* dbeta, void* stream);
int ctclip_geglu_fwd(const void* h, void* u, long long rows, int ld_half, void* stream);
int ctclip_geglu_bwd(const void* h, const void* du, void* dh, long long rows, int ld_half, void* stream);
int ctclip_cast_f32_bf16(const float* x, void* y, long long n, void* stream);
/* n independent 2-D bf16 copies in one launch; desc_dev: device array of n x {src, dst, rows, cols, src_ld, dst_ld} (int64) —
 * re-packs the zero-padded FF operands from the bf16 parameter mirror after an optimiser step */
int ctclip_copy2d_batch_bf16(const long long* desc_dev, int n, void* stream);
/* out[c] += sum_rows x[row][c]  (bias gradients) */
int ctclip_colsum(const float* x, long long rows, int dim, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BERT text tower (reference: ct_clip.py:685-686 -> transformers.BertModel(input_ids, attention_mask)[0];
 * pretrained_model.py:9 injects the BERT-base sized CXR-BERT). The projections and the per-head Q K^T / P V products
 * (and their gradients) run on ctclip_gemm_bf16 (batched mode); these are the kernels between them.
 *   bert_embed_fwd/bwd : BertEmbeddings word + position + token_type(0) gather, and its gradient scatter
 *   bert_softmax_fwd   : BertSelfAttention softmax(scale * scores + key mask) (+ attention-probs dropout)
 *                        scores fp32 [batch*heads][seq][seq], mask int64 [batch][seq] (1 = attend), probs bf16
 *   bert_softmax_bwd   : dscores = scale * P * (dP - rowsum(P dP)), dP = dropout'(dprobs_dropped)
 *   gelu_fwd/bwd       : exact (erf) GELU of BertIntermediate on bf16
 *   dropout_add        : out = dropout(y) (+ resid) (BertSelfOutput / BertOutput hidden dropout; resid NULL = the
 *                        dropout gradient); masks are a stateless hash of (seed, element index), never stored
 *   colsum_bf16        : out[c] += sum_rows x[row][c] on bf16 (bias gradients) */
int ctclip_bert_embed_fwd(const long long* ids, long long tokens, int seq_len, const float* word, const float* pos,
                          const float* type0, int dim, int vocab, float* out, void* stream);
int ctclip_bert_embed_bwd(const long long* ids, long long tokens, int seq_len, const float* dx, int dim, int vocab,
                          long long pad_id /* nn.Embedding padding_idx: no gradient; -1 = none */, float* dword, float* dpos,
                          void* stream);
int ctclip_bert_softmax_fwd(const float* scores, const long long* mask, int batch, int heads, int seq_len, float scale,
                            void* probs, void* probs_dropped, float p_drop, unsigned seed, void* stream);
int ctclip_bert_softmax_bwd(const void* probs, const float* dprobs_dropped, int batch, int heads, int seq_len, float scale,
                            void* dscores, float p_drop, unsigned seed, void* stream);
/* Fused BertSelfAttention core for head dim 64 (HF modeling_bert.py BertSelfAttention.forward: Q K^T / sqrt(d) + key mask ->
 * softmax -> attention-probs dropout -> P V) on the packed [tokens, 3 * hidden] bf16 QKV projection; the scores stay in
 * registers. fwd writes the context [tokens, hidden] bf16 and lse [batch * heads, seq_len] (log2 domain); bwd writes the
 * packed dQKV and needs a [batch * heads, seq_len] fp32 workspace for rowsum(dO * O). */
int ctclip_bert_attn_fwd(const void* qkv, const long long* mask, int batch, int heads, int seq_len, int hidden, void* out,
                         float* lse, float p_drop, unsigned seed, void* stream);
int ctclip_bert_attn_bwd(const void* qkv, const long long* mask, const void* out, const float* lse, const void* dout, int batch,
                         int heads, int seq_len, int hidden, void* dqkv, float* delta_ws, float p_drop, unsigned seed,
                         void* stream);
int ctclip_gelu_fwd(const void* h, void* out, long long n, void* stream);
int ctclip_gelu_bwd(const void* h, const void* dy, void* dh, long long n, void* stream);
int ctclip_dropout_add(const float* y, const float* resid, float* out, long long n, float p_drop, unsigned seed,
                       void* stream);
int ctclip_colsum_bf16(const void* x, long long rows, int dim, long long ld, float* out, void* stream);

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

/* ------------------------------------------------------------------------------------------------
 * Patch embedding front end (ctvit.py:170-171): 'b c (t pt)(h p1)(w p2) -> b t h w (c pt p1 p2)' + LayerNorm(pt*p1*p2),
 * written as the bf16 A operand [tokens][ld_out] of the patch-embed GEMM. video is fp32 [batch][1][frames][height][width].
 * patch_ln_param_grad: dgamma/dbeta (+=) of that LayerNorm derived algebraically from the Linear's W, dW and bias grad. */
int ctclip_patch_ln_fwd(const float* video, int batch, int frames, int height, int width, int pt, int ps,
                        const float* gamma, const float* beta, float eps, void* out, long long ld_out, void* stream);
int ctclip_patch_ln_param_grad(const float* W, const float* dW, const float* s, const float* gamma, const float* beta,
                               float* dgamma, float* dbeta, int n_out, int pdim, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cosine-sim vector quantiser (vector_quantize_pytorch 1.1.2; ctvit.py:187,427).
 * l2norm_rows: y = x / max(|x|,1e-12) (fp32 and/or bf16) + inverse norms.
 * vq_finalize: exact fp32 arg-max from the GEMM's top2_out (n_tiles x float4 per token), embed_n = l2norm(embed) fp32.
 *   stats (may be NULL): [0] += exact dot products evaluated, [1] += whole-tile rescans.
 * vq_gather(_mean): embed[idx] rows / their mean over t (ct_clip.py:724) as fp32 and/or bf16 [batch][hw][dim].
 * pool_bwd: dx[b][t][p][:] = dpool[b][p][:] / t  (mean-over-t backward through the straight-through estimator).
 * vq_ema_*: train-mode EMA statistics (atomics) and the (cluster_size, embed) lerp update with decay 0.8. */
int ctclip_l2norm_rows(const float* x, long long rows, int dim, float* y_f32, void* y_bf16, float* inv_norm, void* stream);
int ctclip_l2norm_bwd(const float* y, const float* inv_norm, const float* g, long long rows, int dim, float* dx, void* stream);
int ctclip_vq_finalize(const void* top2, int n_tiles, int tile_n, const float* x, const float* embed_n, long long rows,
                       int dim, int codes, float margin, int* idx_out, unsigned long long* stats, void* stream);
int ctclip_vq_gather_mean(const float* embed, const int* idx, float* out_f32, void* out_bf16, int batch, int t, int hw,
                          int dim, void* stream);
int ctclip_vq_gather(const float* embed, const int* idx, float* out, long long rows, int dim, void* stream);
int ctclip_pool_bwd(const float* dpool, float* dx, int batch, int t, int hw, int dim, void* stream);
int ctclip_vq_ema_accum(const float* x, const float* inv_norm, const int* idx, long long rows, int dim, float* bins,
                        float* embed_sum, void* stream);
int ctclip_vq_ema_update(float* embed, float* cluster_size, const float* bins, const float* embed_sum, int codes, int dim,
                         float decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Symmetric InfoNCE (ct_clip.py:796,845-878) over the global batch of l2-normalised latents T, I fp32 [B][d].
 * work: fp32 [B*B + 2*B + d]. loss: device scalar. dT/dI: fp32 [rows_local][d], gradients w.r.t. the normalised latents of
 * rows [row0,row0+rows_local) (NULL -> loss only). dtau += this rank's share of dloss/dtemperature. */
int ctclip_clip_loss(const float* T, const float* I, const float* tau, int B, int d, int row0, int rows_local, float* work,
                     float* loss, float* dT, float* dI, float* dtau, void* stream);

/* Zero-shot scoring (ctclip_inference.py:286-336, CTCLIPTrainer.py:356-454): I fp32 [V][d] normalised image latents (each
 * volume encoded ONCE; the reference re-encodes it per pathology), T fp32 [2P][d] normalised prompt latents ordered
 * [present_0, absent_0, present_1, ...]. prob[v][p] = softmax(exp(tau) <I_v, T_2p>, exp(tau) <I_v, T_2p+1>)[0];
 * logits (may be NULL): fp32 [V][P][2]. */
int ctclip_zero_shot_scores(const float* I, const float* T, const float* tau, int V, int P, int d, float* prob, float* logits,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel InfoNCE over NVLink peer memory: the all-gather of the l2-normalised latents fused with
 * the global-batch logits. The reference trains with local-batch InfoNCE under DDP (CTCLIPTrainer.py:213-217 around
 * ct_clip.py:845-878); this entry replaces "NCCL all-gather -> ctclip_clip_loss" of the global-batch extension.
 *
 * Symmetric buffers: every rank allocates one buffer of ctclip_symm_latent_bytes(b_local, d, world) bytes with
 * ctclip_symm_alloc (plain cudaMalloc, zero-filled: the ONE place the library owns device memory, because CUDA IPC can
 * only export cudaMalloc allocations), exports a 64-byte IPC handle (ctclip_symm_export), exchanges handles out of band
 * (torch.distributed all_gather_object) and maps every peer's buffer (ctclip_symm_import; peer access is enabled
 * lazily). host_peer_bufs is a HOST array of `world` device pointers, [rank] = the rank's own buffer.
 *
 * ctclip_clip_loss_allgather: t_hat / i_hat fp32 [b_local][d] are THIS rank's normalised latents. One kernel pushes
 * them into every peer's buffer (16-byte peer stores + st.release.sys of a per-rank flag carrying `step`) and computes
 * each b_local x b_local block of L = exp(tau) T I^T as soon as the two ranks it depends on have arrived
 * (ld.acquire.sys); then the log-sum-exp and gradient kernels of ctclip_clip_loss run on the gathered latents in
 * place. `step` must be the same on all ranks, start at 1 and increase by 1 per call (two buffer parities); all ranks
 * must call collectively, on one stream per rank. A peer that does not arrive within the timeout (default 600 s;
 * CTCLIP_PEER_TIMEOUT_S or ctclip_symm_set_timeout_ms) turns the loss into NaN AND sets bit 0 of *status (device int, may
 * be NULL) instead of hanging; ctclip_adam_step never applies an update whose gradient norm is not finite.
 * work: fp32 [B*B + 2*B + d] with B = world*b_local; outputs as ctclip_clip_loss.
 *
 * ctclip_clip_loss_allgather_emulated: all `world` ranks on ONE device in ONE cooperative launch (bring-up and single-GPU
 * test of the push / flag / wait protocol; mutually waiting kernels must never be separate launches on one GPU).
 * host_bufs: `world` symmetric buffers of this device; t_hat_all / i_hat_all fp32 [world*b_local][d]; per-rank outputs back
 * to back: work_all [world][B*B + 2*B + d], loss_all [world], dT_all / dI_all [world*b_local][d], dtau_all [world]. */
size_t ctclip_symm_latent_bytes(int b_local, int d, int world);
int ctclip_symm_alloc(size_t bytes, void** ptr);
int ctclip_symm_free(void* ptr);
int ctclip_symm_export(void* ptr, unsigned char* handle64);
int ctclip_symm_import(const unsigned char* handle64, void** ptr);
int ctclip_symm_unimport(void* ptr);
int ctclip_symm_set_timeout_ms(unsigned long long ms);
int ctclip_clip_loss_allgather(const float* t_hat, const float* i_hat, const float* tau, int b_local, int d, int rank,
                               int world, void* const* host_peer_bufs, unsigned step, float* work, float* loss, float* dT,
                               float* dI, float* dtau, int* status, void* stream);
int ctclip_clip_loss_allgather_emulated(const float* t_hat_all, const float* i_hat_all, const float* tau, int b_local, int d,
                                        int world, void* const* host_bufs, unsigned step, float* work_all, float* loss_all,
                                        float* dT_all, float* dI_all, float* dtau_all, int* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Continuous position bias (attention.py:229-276; called ctvit.py:317): the 2 -> dim -> dim -> heads MLP with
 * LeakyReLU(0.1) on sign(v)*log(|v|+1) of the relative offsets of an h x w grid, evaluated on the R = (2h-1)(2w-1)
 * DISTINCT offsets (the reference evaluates all (h*w)^2 pairs). table[head][(dy+h-1)*(2w-1) + (dx+w-1)] for the offset
 * (dy, dx) = query - key; the attention kernels gather from it. rowmax (may be NULL): fp32 [heads][h*w] = max over
 * keys of the gathered bias per query position. acts: fp32 workspace [2*R + 2*R*dim] kept for the backward.
 * W0 [dim][2], b0 [dim], W1 [dim][dim], b1 [dim], W2 [heads][dim], b2 [heads] = net.0.0 / net.1.0 / net.2 of the module.
 * Backward: dtable fp32 [heads][R]; work fp32 [2*R*dim]; the six gradients are overwritten. fp32 throughout. */
int ctclip_cpb_table_fwd(int h, int w, int dim, int heads, int log_dist, const float* W0, const float* b0, const float* W1,
                         const float* b1, const float* W2, const float* b2, float* acts, float* table, float* rowmax,
                         void* stream);
int ctclip_cpb_table_bwd(int h, int w, int dim, int heads, const float* W1, const float* W2, const float* acts,
                         const float* dtable, float* work, float* dW0, float* db0, float* dW1, float* db1, float* dW2,
                         float* db2, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Trainer step (CTCLIPTrainer.py:347-353, optimizer.py:10-24) on flat fp32 arenas: sum of squares for the global-norm
 * clip (deterministic: fixed grid of per-block partials added in index order, no atomics — bit-identical on every
 * data-parallel rank), then fused clip + Adam + bf16 shadow refresh + gradient zeroing. A non-finite *norm_sq (NaN / Inf gradients, e.g. a
 * data-parallel peer that timed out) SKIPS the whole update — parameters, moments and gradients stay as they are — and
 * increments *skipped (device int, may be NULL); the host raises on it (CTClipTrainStep.raise_if_skipped). */
int ctclip_sumsq(const float* g, long long n, float* out, float* workspace /* ctclip_workspace_bytes("sumsq") */, void* stream);
int ctclip_adam_step(float* p, float* g, float* m, float* v, void* bf16_shadow, long long n, float lr, float beta1,
                     float beta2, float eps, int step, const float* norm_sq, float max_norm, int zero_grad, int* skipped,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * data_prep volume normalisation (preprocess_train.py:99-109, resize_array :31-42, data.py:155-190).
 *   in_is_i16=1: raw scan, int16, logical (D,H,W) addressed through element strides (a NIfTI (H,W,N) array has
 *   stride_d=1, stride_w=N, stride_h=W*N); HU = slope*x+intercept in float64, clip [-1000,1000], /1000, float32.
 *   in_is_i16=0: float32 volume, resample only (ctpa_report/vqa_meditron.py:170-175 direct-to-target variant).
 *   (oD,oH,oW): trilinear target grid (align_corners=False). (tD,tH,tW) > 0: destination volume receiving the
 *   centre-cropped / pad_value-padded result (data.py:155-189); 0 -> destination == target grid. */
typedef struct ctclip_prep_desc {
  const void* in; float* out;
  int batch; int in_is_i16;
  double slope, intercept;
  int D, H, W;
  long long stride_d, stride_h, stride_w, stride_batch;
  int oD, oH, oW;
  int tD, tH, tW;
  float pad_value;
  float* lut_workspace; /* device scratch of >= 8192 floats for the exact HU table (int16 input); NULL -> per-voxel fp64 */
  int force_generic;    /* 1: skip the depth-marching fast path for int16 (H,W,N) scans (test hook; results are identical) */
  /* fp32 input only — the value arithmetic of the two DataLoaders, fused around the resample (float32, one rounding per
   * numpy operation):
   *   pre_op 2: CTReportDatasetinfer.nii_img_to_tensor, ct_clip/data_inference.py:81-85: (clip(x*1000, -1000, 200) + 400) / 600
   *   pre_op 3: CTReportDataset.npz_img_to_tensor, ct_clip/data.py:138: (float)slope * x + (float)intercept
   *   post_op 1: ct_clip/data.py:150-152, after the resample: clip(v, -1000, 1000) / 1000 */
  int pre_op, post_op;
} ctclip_prep_desc;
int ctclip_prep_resample(const ctclip_prep_desc* d, void* stream);
/* 12-bit transfer format of raw scans (the loaders of data_prep/preprocess_train.py:60-109 and ct_clip/data.py:114-150 hold
 * int16 / float arrays; CT voxels carry 12 significant bits): `in` holds n_voxels * 3 / 2 bytes, two voxels v = raw + offset
 * (clamped to [0, 4095]) per three bytes, little endian (b0 = v0 & 0xff, b1 = v0 >> 8 | (v1 & 0xf) << 4, b2 = v1 >> 4);
 * `out` receives the n_voxels int16 values v - offset that ctclip_prep_resample reads (in_is_i16). n_voxels % 16 == 0.
 * Host-side packer: ctpa_clip_b200/data_prep/pack12.py. Lossless for the path whenever [-offset, 4095 - offset] covers the
 * pre-image of the HU window [-1000, 1000] process_file clips to. */
int ctclip_unpack12(const void* in, void* out, long long n_voxels, int offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCLIP_B200_H_ */
