"""Tensor-level wrappers over the C-ABI: take torch CUDA tensors, pass raw pointers + the current stream.

PyTorch is only the allocator / stream provider here; all arithmetic happens in libctclip_sm100.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


_PROF = None  # {name: [(start_event, end_event, flops), ...]} while bench.py's instrumented step runs


def profile_begin():
    global _PROF
    _PROF = {}


def profile_end():
    """returns {op name: {"ms": summed device time, "n": launches, "flops": algorithmic flops}}"""
    global _PROF
    prof, _PROF = _PROF, None
    torch.cuda.synchronize()
    out = {}
    for name, recs in prof.items():
        out[name] = {"ms": sum(a.elapsed_time(b) for a, b, _ in recs), "n": len(recs), "flops": sum(f for _, _, f in recs)}
    return out


class _Span:
    def __init__(self, name, flops=0.0):
        self.name, self.flops = name, flops

    def __enter__(self):
        if _PROF is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PROF is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _PROF.setdefault(self.name, []).append((self.a, b, self.flops))
        return False


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> C.c_void_p:
    """the current torch stream of the current device as a cudaStream_t. torch.cuda.current_stream() builds a Stream object
    through three layers of Python (14 us per call, 600 calls per step); the raw query is one C call."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.CtclipError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.CtclipError(f"{name}: expected dtype {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_t: bool = False, b_t: bool = False, out: torch.Tensor | None = None,
         out_dtype=torch.bfloat16, bias: torch.Tensor | None = None, resid: torch.Tensor | None = None,
         alpha: float = 1.0, accumulate: bool = False, splits: int = 1, tag: str = "", flops: float | None = None) -> torch.Tensor:
    """C[M,N] = alpha * op(A) op(B)^T (+bias) (+resid)   bf16 operands, fp32 accumulation (tcgen05).

    a_t=False: `a` is [M,K] row-major;  a_t=True: `a` is [K,M] row-major (A^T stored).
    b_t=False: `b` is [N,K] row-major;  b_t=True: `b` is [K,N] row-major.
    accumulate=True: atomically add into fp32 `out` (split-K allowed, splits=0 -> auto).
    Row strides may exceed the logical width (padded operands) but the last dim must be contiguous.
    """
    _req(a, torch.bfloat16, "gemm.a")
    _req(b, torch.bfloat16, "gemm.b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_t:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_t:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    if K != Kb:
        raise _lib.CtclipError(f"gemm: K mismatch {K} vs {Kb}")
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert out.dim() == 2 and out.shape[0] == M and out.shape[1] == N and out.stride(1) == 1
    d = _lib.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = a.data_ptr(), a.stride(0), int(a_t)
    d.B, d.ldb, d.b_mn_major = b.data_ptr(), b.stride(0), int(b_t)
    d.C, d.ldc = out.data_ptr(), out.stride(0)
    if out.dtype == torch.float32:
        d.c_is_f32 = 1
    elif out.dtype == torch.bfloat16:
        d.c_is_f32 = 0
    else:
        raise _lib.CtclipError("gemm: out must be bf16 or fp32")
    if bias is not None:
        _req(bias, torch.float32, "gemm.bias")
        d.bias = bias.data_ptr()
    if resid is not None:
        _req(resid, torch.float32, "gemm.resid")
        assert resid.shape == out.shape and resid.stride(1) == 1
        d.resid, d.ldr = resid.data_ptr(), resid.stride(0)
    d.alpha = alpha
    d.atomic = int(accumulate)
    d.splits = splits
    with _Span("gemm:" + (tag or f"{M}x{N}x{K}"), 2.0 * M * N * K if flops is None else flops):
        _lib.check(_lib.lib().ctclip_gemm_bf16(C.byref(d), _stream()), "ctclip_gemm_bf16")
    return out


def unpack12(packed: torch.Tensor, out: torch.Tensor, n_voxels: int, offset: int = 1024):
    """12-bit packed raw scans (uint8, n_voxels * 3 / 2 bytes) -> int16 `out` (n_voxels elements); see data_prep/pack12.py"""
    _req(packed, torch.uint8, "unpack12.packed")
    _req(out, torch.int16, "unpack12.out")
    assert packed.is_contiguous() and out.is_contiguous() and packed.numel() * 2 == n_voxels * 3 and out.numel() == n_voxels
    _call("ctclip_unpack12", _ptr(packed), _ptr(out), _ll(n_voxels), int(offset), _stream())
    return out


def gemm_geglu(a: torch.Tensor, w: torch.Tensor, *, tag: str = ""):
    """FeedForward's first Linear with GEGLU in the epilogue (attention.py:39-48): a [M, K] bf16, w [2 Nh, K] bf16 =
    [x rows | gate rows]. Returns (h [M, 2 Nh] bf16 = [x | gate], kept for the backward, u [M, Nh] bf16 = x * gelu(gate)) —
    bit-identical to geglu_fwd(gemm(a, w)) without re-reading h."""
    _req(a, torch.bfloat16, "gemm_geglu.a")
    _req(w, torch.bfloat16, "gemm_geglu.w")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if N % 16:
        raise _lib.CtclipError("gemm_geglu: 2 Nh must be a multiple of 16")
    h = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    u = torch.empty((M, N // 2), device=a.device, dtype=torch.bfloat16)
    d = _lib.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = a.data_ptr(), a.stride(0), 0
    d.B, d.ldb, d.b_mn_major = w.data_ptr(), w.stride(0), 0
    d.C, d.ldc, d.c_is_f32 = h.data_ptr(), h.stride(0), 0
    d.alpha, d.splits = 1.0, 1
    d.geglu_u, d.ld_u = u.data_ptr(), u.stride(0)
    with _Span("gemm:" + (tag or f"{M}x{N}x{K}+geglu"), 2.0 * M * N * K):
        _lib.check(_lib.lib().ctclip_gemm_bf16(C.byref(d), _stream()), "ctclip_gemm_bf16(geglu)")
    return h, u


def _call(name: str, *args):
    fn = getattr(_lib.lib(), name)
    with _Span(name.replace("ctclip_", "")):
        _lib.check(fn(*args), name)


def _f(x: float):
    return C.c_float(x)


def _ll(x: int):
    return C.c_longlong(x)


def layernorm_fwd(x, gamma, beta=None, *, eps=1e-5, want_bf16=True, want_raw_bf16=False, want_f32=False):
    """x fp32 [rows, dim] -> (y_bf16 | None, raw_bf16 | None, y_f32 | None)"""
    _req(x, torch.float32, "layernorm_fwd.x")
    rows, dim = x.shape
    assert x.is_contiguous() and gamma.dtype == torch.float32
    y = torch.empty((rows, dim), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    raw = torch.empty((rows, dim), device=x.device, dtype=torch.bfloat16) if want_raw_bf16 else None
    yf = torch.empty((rows, dim), device=x.device, dtype=torch.float32) if want_f32 else None
    _call("ctclip_layernorm_fwd", _ptr(x), _ll(rows), dim, _ptr(gamma), _ptr(beta), _f(eps), _ptr(y), _ptr(raw),
          _ptr(yf), _stream())
    return y, raw, yf


def layernorm_bwd(dy, x, gamma, *, eps=1e-5, add_in=None, dgamma=None, dbeta=None, want_bf16=False, out=None):
    """returns (dx fp32, dx_bf16 | None); dgamma / dbeta are accumulated in place"""
    if dy.dtype not in (torch.float32, torch.bfloat16):
        raise _lib.CtclipError("layernorm_bwd.dy: expected fp32 or bf16")
    _req(x, torch.float32, "layernorm_bwd.x")
    rows, dim = x.shape
    assert dy.is_cuda and dy.is_contiguous() and x.is_contiguous() and dy.shape == x.shape
    dx = out if out is not None else torch.empty_like(x)
    dxb = torch.empty((rows, dim), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    _call("ctclip_layernorm_bwd", _ptr(dy), int(dy.dtype == torch.bfloat16), _ptr(x), _ll(rows), dim, _ptr(gamma), _f(eps),
          _ptr(add_in), _ptr(dx), _ptr(dxb), _ptr(dgamma), _ptr(dbeta), _stream())
    return dx, dxb


def geglu_fwd(h):
    _req(h, torch.bfloat16, "geglu_fwd.h")
    rows, two = h.shape
    assert h.is_contiguous() and two % 16 == 0
    u = torch.empty((rows, two // 2), device=h.device, dtype=torch.bfloat16)
    _call("ctclip_geglu_fwd", _ptr(h), _ptr(u), _ll(rows), two // 2, _stream())
    return u


def geglu_bwd(h, du):
    _req(h, torch.bfloat16, "geglu_bwd.h")
    _req(du, torch.bfloat16, "geglu_bwd.du")
    rows, two = h.shape
    assert h.is_contiguous() and du.is_contiguous() and du.shape == (rows, two // 2)
    dh = torch.empty_like(h)
    _call("ctclip_geglu_bwd", _ptr(h), _ptr(du), _ptr(dh), _ll(rows), two // 2, _stream())
    return dh


def cast_bf16(x, out=None):
    _req(x, torch.float32, "cast_bf16.x")
    assert x.is_contiguous()
    y = out if out is not None else torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    assert y.dtype == torch.bfloat16 and y.numel() == x.numel() and y.is_contiguous()
    _call("ctclip_cast_f32_bf16", _ptr(x), _ptr(y), _ll(x.numel()), _stream())
    return y


def peg_fwd(x, w27, bias, grid, temporal: bool):
    """x fp32 [b*t*h*w, dim] canonical -> x + dwconv(x) + bias"""
    _req(x, torch.float32, "peg_fwd.x")
    b, t, h, w = grid
    y = torch.empty_like(x)
    _call("ctclip_peg_fwd", _ptr(x), _ptr(y), _ptr(w27), _ptr(bias), b, t, h, w, x.shape[-1], int(temporal), _stream())
    return y


def peg_bwd_data(dy, w27, grid, temporal: bool, want_bf16=False):
    _req(dy, torch.float32, "peg_bwd_data.dy")
    b, t, h, w = grid
    dx = torch.empty_like(dy)
    dxb = torch.empty(dy.shape, device=dy.device, dtype=torch.bfloat16) if want_bf16 else None
    _call("ctclip_peg_bwd_data", _ptr(dy), _ptr(dx), _ptr(dxb), _ptr(w27), b, t, h, w, dy.shape[-1], int(temporal),
          _stream())
    return dx, dxb


def peg_bwd_weight(x, dy, dw27, dbias, grid, temporal: bool):
    b, t, h, w = grid
    _call("ctclip_peg_bwd_weight", _ptr(x), _ptr(dy), _ptr(dw27), _ptr(dbias), b, t, h, w, x.shape[-1], int(temporal),
          _stream())


def _attn_desc(q, kv, grid, heads, temporal, q_scale, k_scale, bias_table, bias_rowmax):
    b, t, h, w = grid
    d = _lib.AttnDesc()
    d.batch, d.t, d.h, d.w, d.heads, d.dim_head, d.temporal = b, t, h, w, heads, 32, int(temporal)
    d.q, d.ldq = q.data_ptr(), q.stride(0)
    d.kv, d.ldkv = kv.data_ptr(), kv.stride(0)
    d.q_scale, d.k_scale = q_scale.data_ptr(), k_scale.data_ptr()
    if bias_table is not None and not temporal:
        d.bias_table, d.bias_rowmax = bias_table.data_ptr(), bias_rowmax.data_ptr()
    return d


def attn_fwd(q, kv, grid, heads, temporal, q_scale, k_scale, bias_table=None, bias_rowmax=None):
    """q bf16 [tokens, heads*32], kv bf16 [tokens, 2*heads*32] -> (o bf16 [tokens, heads*32], lse fp32 [tokens, heads])"""
    _req(q, torch.bfloat16, "attn_fwd.q")
    _req(kv, torch.bfloat16, "attn_fwd.kv")
    tokens = q.shape[0]
    o = torch.empty((tokens, heads * 32), device=q.device, dtype=torch.bfloat16)
    lse = torch.empty((tokens, heads), device=q.device, dtype=torch.float32)
    d = _attn_desc(q, kv, grid, heads, temporal, q_scale, k_scale, bias_table, bias_rowmax)
    d.o, d.ldo, d.lse = o.data_ptr(), o.stride(0), lse.data_ptr()
    with _Span("attn_fwd:" + ("temporal" if temporal else "spatial")):
        _lib.check(_lib.lib().ctclip_attn_fwd(C.byref(d), _stream()), "ctclip_attn_fwd")
    return o, lse


def attn_bwd(q, kv, o, lse, d_o, grid, heads, temporal, q_scale, k_scale, dq_scale, dk_scale, bias_table=None,
             bias_rowmax=None, dbias_table=None):
    """returns (dq bf16, dkv bf16); dq_scale / dk_scale / dbias_table accumulated in place"""
    tokens = q.shape[0]
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    d = _attn_desc(q, kv, grid, heads, temporal, q_scale, k_scale, bias_table, bias_rowmax)
    d.o, d.ldo, d.lse = o.data_ptr(), o.stride(0), lse.data_ptr()
    d.d_o = d_o.data_ptr()
    assert d_o.stride(0) == o.stride(0) and dq.stride(0) == q.stride(0) and dkv.stride(0) == kv.stride(0)
    d.dq, d.dkv = dq.data_ptr(), dkv.data_ptr()
    d.dq_scale, d.dk_scale = dq_scale.data_ptr(), dk_scale.data_ptr()
    if dbias_table is not None and not temporal:
        d.dbias_table = dbias_table.data_ptr()
    with _Span("attn_bwd:" + ("temporal" if temporal else "spatial")):
        _lib.check(_lib.lib().ctclip_attn_bwd(C.byref(d), _stream()), "ctclip_attn_bwd")
    return dq, dkv


def colsum(x, out):
    _req(x, torch.float32, "colsum.x")
    assert x.is_contiguous()
    _call("ctclip_colsum", _ptr(x), _ll(x.shape[0]), x.shape[1], _ptr(out), _stream())


def gemm_top2(a, b):
    """VQ assignment GEMM: returns (top2 float32 [M, n_tiles, 4], tile_n) without materialising the score matrix"""
    _req(a, torch.bfloat16, "gemm_top2.a")
    _req(b, torch.bfloat16, "gemm_top2.b")
    M, K = a.shape
    N = b.shape[0]
    tile_n = int(_lib.lib().ctclip_gemm_tile_n(N))
    n_tiles = (N + tile_n - 1) // tile_n
    top2 = torch.empty((M, n_tiles, 4), device=a.device, dtype=torch.float32)
    d = _lib.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = a.data_ptr(), a.stride(0), 0
    d.B, d.ldb, d.b_mn_major = b.data_ptr(), b.stride(0), 0
    d.C, d.ldc, d.c_is_f32 = 0, N, 1
    d.alpha, d.atomic, d.splits = 1.0, 0, 1
    d.top2_out = top2.data_ptr()
    with _Span("gemm:vq_top2", 2.0 * M * N * K):
        _lib.check(_lib.lib().ctclip_gemm_bf16(C.byref(d), _stream()), "ctclip_gemm_bf16(top2)")
    return top2, tile_n


def patch_ln_fwd(video, gamma, beta, pt, ps, eps=1e-5):
    """video fp32 [b,1,f,h,w] -> bf16 [tokens, pdim_padded] LayerNorm'd patches in (pt p1 p2) order"""
    _req(video, torch.float32, "patch_ln_fwd.video")
    assert video.is_contiguous() and video.dim() == 5 and video.shape[1] == 1
    b, _, f, hh, ww = video.shape
    pdim = pt * ps * ps
    ld = (pdim + 7) // 8 * 8
    tokens = b * (f // pt) * (hh // ps) * (ww // ps)
    alloc = torch.zeros if ld != pdim else torch.empty
    out = alloc((tokens, ld), device=video.device, dtype=torch.bfloat16)
    _call("ctclip_patch_ln_fwd", _ptr(video), b, f, hh, ww, pt, ps, _ptr(gamma), _ptr(beta), _f(eps), _ptr(out), _ll(ld),
          _stream())
    return out


def patch_ln_param_grad(W, dW, s, gamma, beta, dgamma, dbeta):
    n_out, pdim = W.shape
    assert W.is_contiguous() and dW.is_contiguous() and dW.shape == W.shape
    _call("ctclip_patch_ln_param_grad", _ptr(W), _ptr(dW), _ptr(s), _ptr(gamma), _ptr(beta), _ptr(dgamma), _ptr(dbeta),
          n_out, pdim, _stream())


def l2norm_rows(x, want_f32=False, want_bf16=False):
    """returns (y_f32 | None, y_bf16 | None, inv_norm)"""
    _req(x, torch.float32, "l2norm_rows.x")
    assert x.is_contiguous() and x.dim() == 2
    rows, dim = x.shape
    yf = torch.empty_like(x) if want_f32 else None
    yb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    inv = torch.empty(rows, device=x.device, dtype=torch.float32)
    _call("ctclip_l2norm_rows", _ptr(x), _ll(rows), dim, _ptr(yf), _ptr(yb), _ptr(inv), _stream())
    return yf, yb, inv


def l2norm_bwd(y, inv_norm, g):
    dx = torch.empty_like(y)
    _call("ctclip_l2norm_bwd", _ptr(y), _ptr(inv_norm), _ptr(g), _ll(y.shape[0]), y.shape[1], _ptr(dx), _stream())
    return dx


VQ_MARGIN = 8e-3  # >= 2 * 2^-8: twice the worst-case error of a unit-vector dot product with bf16-rounded operands


def vq_finalize(top2, tile_n, x, embed_n, margin=VQ_MARGIN, stats=None):
    rows, dim = x.shape
    codes = embed_n.shape[0]
    idx = torch.empty(rows, device=x.device, dtype=torch.int32)
    _call("ctclip_vq_finalize", _ptr(top2), top2.shape[1], tile_n, _ptr(x), _ptr(embed_n), _ll(rows), dim, codes,
          _f(margin), _ptr(idx), _ptr(stats), _stream())
    return idx


def vq_gather_mean(embed, idx, batch, t, hw, want_bf16=True):
    dim = embed.shape[-1]
    out = torch.empty((batch, hw * dim), device=embed.device, dtype=torch.float32)
    outb = torch.empty((batch, hw * dim), device=embed.device, dtype=torch.bfloat16) if want_bf16 else None
    _call("ctclip_vq_gather_mean", _ptr(embed), _ptr(idx), _ptr(out), _ptr(outb), batch, t, hw, dim, _stream())
    return out, outb


def vq_gather(embed, idx):
    dim = embed.shape[-1]
    out = torch.empty((idx.numel(), dim), device=embed.device, dtype=torch.float32)
    _call("ctclip_vq_gather", _ptr(embed), _ptr(idx), _ptr(out), _ll(idx.numel()), dim, _stream())
    return out


def pool_bwd(dpool, batch, t, hw, dim):
    dx = torch.empty((batch * t * hw, dim), device=dpool.device, dtype=torch.float32)
    _call("ctclip_pool_bwd", _ptr(dpool), _ptr(dx), batch, t, hw, dim, _stream())
    return dx


def vq_ema(embed, cluster_size, x, inv_norm, idx, decay=0.8):
    """in-place train-mode EMA update of (embed [C,D], cluster_size [C]); returns (bins, embed_sum) for cross-rank reduction"""
    codes, dim = embed.shape
    # ONE buffer [bins | embed_sum]: a single collective reduces both across data-parallel ranks (bins.storage is the buffer)
    buf = torch.zeros(codes * (dim + 1), device=embed.device, dtype=torch.float32)
    bins, esum = buf[:codes], buf[codes:].view(codes, dim)
    _call("ctclip_vq_ema_accum", _ptr(x), _ptr(inv_norm), _ptr(idx), _ll(x.shape[0]), dim, _ptr(bins), _ptr(esum), _stream())
    return bins, esum


def vq_ema_update(embed, cluster_size, bins, esum, decay=0.8):
    codes, dim = embed.shape
    _call("ctclip_vq_ema_update", _ptr(embed), _ptr(cluster_size), _ptr(bins), _ptr(esum), codes, dim, _f(decay), _stream())


def clip_loss(T, I, tau, row0, rows_local, want_grad=True):
    """T, I fp32 [B, d] normalised global latents; returns (loss scalar tensor, dT, dI, dtau) for the local rows"""
    B, d = T.shape
    work = torch.empty(B * B + 2 * B + d, device=T.device, dtype=torch.float32)
    loss = torch.zeros((), device=T.device, dtype=torch.float32)
    dT = torch.empty((rows_local, d), device=T.device, dtype=torch.float32) if want_grad else None
    dI = torch.empty((rows_local, d), device=T.device, dtype=torch.float32) if want_grad else None
    dtau = torch.zeros((), device=T.device, dtype=torch.float32) if want_grad else None
    _call("ctclip_clip_loss", _ptr(T), _ptr(I), _ptr(tau), B, d, row0, rows_local, _ptr(work), _ptr(loss), _ptr(dT),
          _ptr(dI), _ptr(dtau), _stream())
    return loss, dT, dI, dtau


def clip_loss_allgather(t_hat, i_hat, tau, rank, world, peer_table, step, status=None):
    """Peer-memory latent exchange fused with the global-batch logits (csrc/symm.cu), then lse + gradient of the local rows.
    t_hat, i_hat fp32 [b, d] normalised LOCAL latents; peer_table: ctypes array of `world` symmetric-buffer pointers;
    status: int32 device scalar whose bit 0 is raised when a peer did not arrive within the timeout (loss = NaN)."""
    b, d = t_hat.shape
    B = world * b
    work = torch.empty(B * B + 2 * B + d, device=t_hat.device, dtype=torch.float32)
    loss = torch.zeros((), device=t_hat.device, dtype=torch.float32)
    dT = torch.empty((b, d), device=t_hat.device, dtype=torch.float32)
    dI = torch.empty((b, d), device=t_hat.device, dtype=torch.float32)
    dtau = torch.zeros((), device=t_hat.device, dtype=torch.float32)
    _call("ctclip_clip_loss_allgather", _ptr(t_hat), _ptr(i_hat), _ptr(tau), b, d, rank, world, peer_table, C.c_uint(step),
          _ptr(work), _ptr(loss), _ptr(dT), _ptr(dI), _ptr(dtau), _ptr(status), _stream())
    return loss, dT, dI, dtau


def clip_loss_allgather_emulated(t_hat_all, i_hat_all, tau, world, bufs, step, status=None):
    """every rank of the peer-memory exchange on ONE device in ONE cooperative launch (`ctclip_clip_loss_allgather_emulated`):
    t_hat_all / i_hat_all fp32 [world*b, d]; bufs: ctypes array of `world` symmetric buffers on this device.
    Returns per-rank (loss [world], dT [world*b, d], dI [world*b, d], dtau [world])."""
    B, d = t_hat_all.shape
    b = B // world
    dev = t_hat_all.device
    work = torch.empty(world * (B * B + 2 * B + d), device=dev, dtype=torch.float32)
    loss = torch.zeros(world, device=dev, dtype=torch.float32)
    dT = torch.empty((B, d), device=dev, dtype=torch.float32)
    dI = torch.empty((B, d), device=dev, dtype=torch.float32)
    dtau = torch.zeros(world, device=dev, dtype=torch.float32)
    _call("ctclip_clip_loss_allgather_emulated", _ptr(t_hat_all), _ptr(i_hat_all), _ptr(tau), b, d, world, bufs,
          C.c_uint(step), _ptr(work), _ptr(loss), _ptr(dT), _ptr(dI), _ptr(dtau), _ptr(status), _stream())
    return loss, dT, dI, dtau


def cpb_table_fwd(h, w, W0, b0, W1, b1, W2, b2, log_dist=True, want_rowmax=True):
    """continuous-position-bias MLP on the (2h-1)(2w-1) distinct offsets -> (table [heads, R], rowmax [heads, h*w], acts)"""
    dim, heads = W1.shape[0], W2.shape[0]
    for t in (W0, b0, W1, b1, W2, b2):
        _req(t, torch.float32, "cpb_table_fwd")
    R = (2 * h - 1) * (2 * w - 1)
    dev = W0.device
    acts = torch.empty(2 * R + 2 * R * dim, device=dev, dtype=torch.float32)
    table = torch.empty((heads, R), device=dev, dtype=torch.float32)
    rowmax = torch.empty((heads, h * w), device=dev, dtype=torch.float32) if want_rowmax else None
    _call("ctclip_cpb_table_fwd", h, w, dim, heads, int(log_dist), _ptr(W0.contiguous()), _ptr(b0.contiguous()),
          _ptr(W1.contiguous()), _ptr(b1.contiguous()), _ptr(W2.contiguous()), _ptr(b2.contiguous()), _ptr(acts), _ptr(table),
          _ptr(rowmax), _stream())
    return table, rowmax, acts


def cpb_table_bwd(h, w, W1, W2, acts, dtable):
    """gradients of the CPB MLP parameters from dtable [heads, R]: (dW0, db0, dW1, db1, dW2, db2)"""
    dim, heads = W1.shape[0], W2.shape[0]
    R = (2 * h - 1) * (2 * w - 1)
    dev = W1.device
    _req(dtable, torch.float32, "cpb_table_bwd")
    work = torch.empty(2 * R * dim, device=dev, dtype=torch.float32)
    dW0 = torch.empty((dim, 2), device=dev, dtype=torch.float32)
    db0 = torch.empty(dim, device=dev, dtype=torch.float32)
    dW1 = torch.empty((dim, dim), device=dev, dtype=torch.float32)
    db1 = torch.empty(dim, device=dev, dtype=torch.float32)
    dW2 = torch.empty((heads, dim), device=dev, dtype=torch.float32)
    db2 = torch.empty(heads, device=dev, dtype=torch.float32)
    _call("ctclip_cpb_table_bwd", h, w, dim, heads, _ptr(W1.contiguous()), _ptr(W2.contiguous()), _ptr(acts),
          _ptr(dtable.contiguous()), _ptr(work), _ptr(dW0), _ptr(db0), _ptr(dW1), _ptr(db1), _ptr(dW2), _ptr(db2), _stream())
    return dW0, db0, dW1, db1, dW2, db2


def zero_shot_scores(i_hat, t_hat, tau, want_logits=False):
    """i_hat fp32 [V, d], t_hat fp32 [2P, d] (present/absent pairs) -> prob[present] [V, P] (+ logits [V, P, 2])"""
    _req(i_hat, torch.float32, "zero_shot_scores")
    _req(t_hat, torch.float32, "zero_shot_scores")
    V, d = i_hat.shape
    P = t_hat.shape[0] // 2
    prob = torch.empty((V, P), device=i_hat.device, dtype=torch.float32)
    logits = torch.empty((V, P, 2), device=i_hat.device, dtype=torch.float32) if want_logits else None
    _call("ctclip_zero_shot_scores", _ptr(i_hat.contiguous()), _ptr(t_hat.contiguous()), _ptr(tau), V, P, d, _ptr(prob),
          _ptr(logits), _stream())
    return (prob, logits) if want_logits else prob


_SUMSQ_WS = {}


def sumsq(g, out, workspace=None):
    """out[0] += sum(g^2), deterministic (no atomics); workspace: 1024 floats of scratch (one cached per device if None)"""
    if workspace is None:
        workspace = _SUMSQ_WS.get(g.device)
        if workspace is None:
            workspace = _SUMSQ_WS[g.device] = torch.empty(1024, device=g.device, dtype=torch.float32)
    _call("ctclip_sumsq", _ptr(g), _ll(g.numel()), _ptr(out), _ptr(workspace), _stream())


def adam_step(p, g, m, v, shadow, lr, beta1, beta2, eps, step, norm_sq=None, max_norm=0.0, zero_grad=True, skipped=None):
    """skipped: int32 device scalar, incremented (and the whole update skipped) when norm_sq is not finite"""
    _call("ctclip_adam_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(shadow), _ll(p.numel()), _f(lr), _f(beta1),
          _f(beta2), _f(eps), step, _ptr(norm_sq), _f(max_norm), int(zero_grad), _ptr(skipped), _stream())


PREP_PRE_OPS = {None: 0, "infer_window": 2, "affine": 3}
PREP_POST_OPS = {None: 0, "clip_div": 1}


_PREP_LUT = {}


def prep_resample(inp, out_grid, *, hu=None, layout="dhw", target=None, pad_value=-1.0, force_generic=False,
                  pre_op=None, post_op=None, out=None):
    """Trilinear resample (align_corners=False) of a batch of volumes on the GPU.
    layout "dhw": inp is [b, D, H, W] (fp32, or int16 with hu=(slope, intercept));
    layout "hwn": inp is [b, H, W, N] as stored in a NIfTI array (depth contiguous).
    out_grid = (oD, oH, oW) resampled size; target = (tD, tH, tW) -> centre crop / pad(pad_value) window.
    fp32 input: pre_op "affine" (with hu=(slope, intercept), data.py:138) or "infer_window" (data_inference.py:81-85) is applied
    to every voxel before, post_op "clip_div" (data.py:150-152) after the resample — the DataLoaders' float32 arithmetic."""
    assert inp.is_cuda and inp.is_contiguous() and inp.dim() == 4
    d = _lib.PrepDesc()
    b = inp.shape[0]
    if layout == "dhw":
        D, H, W = inp.shape[1:]
        sd, sh, sw = H * W, W, 1
    elif layout == "hwn":
        H, W, D = inp.shape[1:]
        sd, sh, sw = 1, W * D, D
    else:
        raise _lib.CtclipError("prep_resample: unknown layout")
    if inp.dtype == torch.int16:
        if hu is None:
            raise _lib.CtclipError("prep_resample: int16 input needs hu=(slope, intercept)")
        d.in_is_i16, d.slope, d.intercept = 1, float(hu[0]), float(hu[1])
        d.pre_op, d.post_op = PREP_PRE_OPS[pre_op], PREP_POST_OPS[post_op]    # rejected by the library for int16
    elif inp.dtype == torch.float32:
        d.in_is_i16 = 0
        d.pre_op, d.post_op = PREP_PRE_OPS[pre_op], PREP_POST_OPS[post_op]
        if pre_op == "affine":
            d.slope, d.intercept = float(hu[0]), float(hu[1])
    else:
        raise _lib.CtclipError("prep_resample: input must be int16 or float32")
    tgt = tuple(target) if target is not None else tuple(out_grid)
    if out is None:
        out = torch.empty((b, *tgt), device=inp.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != b * tgt[0] * tgt[1] * tgt[2] or not out.is_cuda:
        raise _lib.CtclipError("prep_resample: out must be a contiguous fp32 CUDA tensor of the destination size")
    d.in_, d.out, d.batch = inp.data_ptr(), out.data_ptr(), b
    d.D, d.H, d.W = D, H, W
    d.stride_d, d.stride_h, d.stride_w, d.stride_batch = sd, sh, sw, D * H * W
    d.oD, d.oH, d.oW = (int(v) for v in out_grid)
    d.tD, d.tH, d.tW = (int(v) for v in tgt)
    d.pad_value = pad_value
    d.force_generic = int(force_generic)
    if inp.dtype == torch.int16:
        key = (inp.device, torch.cuda.current_stream().cuda_stream)     # one table per stream: concurrent calls never share it
        lut = _PREP_LUT.get(key)
        if lut is None:
            lut = _PREP_LUT[key] = torch.empty(8192, device=inp.device, dtype=torch.float32)
        d.lut_workspace = lut.data_ptr()
    _call("ctclip_prep_resample", C.byref(d), _stream())
    return out


# ------------------------------------------------------------------------------------------------ BERT text tower
def gemm_batched(a_ptr, lda, a_t, a_sh, a_sb, b_ptr, ldb, b_t, b_sh, b_sb, out, ldc, c_sh, c_sb, M, N, K, H, Bn,
                 alpha=1.0, tag="bmm"):
    """H*Bn independent products C_z[M,N] = alpha * op(A_z) op(B_z)^T on the tcgen05 GEMM (batched mode).
    Operands are raw device addresses (int) of bf16 data + leading dimensions / (head, batch) strides in elements; `out` is a
    bf16 or fp32 tensor that owns the memory written (its data_ptr is the C of problem (0, 0))."""
    d = _lib.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = a_ptr, lda, int(a_t)
    d.B, d.ldb, d.b_mn_major = b_ptr, ldb, int(b_t)
    d.C, d.ldc = out if isinstance(out, int) else out.data_ptr(), ldc
    d.c_is_f32 = 1
    d.alpha, d.atomic, d.splits = alpha, 0, 1
    d.batch_h, d.batch_b = H, Bn
    d.a_stride_h, d.a_stride_b = a_sh, a_sb
    d.b_stride_h, d.b_stride_b = b_sh, b_sb
    d.c_stride_h, d.c_stride_b = c_sh, c_sb
    return d


def run_gemm_desc(d, c_is_f32: bool, tag: str, flops: float):
    d.c_is_f32 = int(c_is_f32)
    with _Span("gemm:" + tag, flops):
        _lib.check(_lib.lib().ctclip_gemm_bf16(C.byref(d), _stream()), "ctclip_gemm_bf16(batched)")


def bert_embed_fwd(ids, seq_len, word, pos, type0):
    tokens, dim = ids.numel(), word.shape[1]
    out = torch.empty((tokens, dim), device=word.device, dtype=torch.float32)
    _call("ctclip_bert_embed_fwd", _ptr(ids), _ll(tokens), seq_len, _ptr(word), _ptr(pos), _ptr(type0), dim, word.shape[0],
          _ptr(out), _stream())
    return out


def bert_embed_bwd(ids, seq_len, dx, dword, dpos, pad_id=-1):
    _call("ctclip_bert_embed_bwd", _ptr(ids), _ll(ids.numel()), seq_len, _ptr(dx), dx.shape[1], dword.shape[0],
          _ll(-1 if pad_id is None else pad_id), _ptr(dword), _ptr(dpos), _stream())


def bert_softmax_fwd(scores, mask, batch, heads, seq_len, scale, p_drop=0.0, seed=0):
    probs = torch.empty(scores.shape, device=scores.device, dtype=torch.bfloat16)
    dropped = torch.empty_like(probs) if p_drop > 0 else None
    _call("ctclip_bert_softmax_fwd", _ptr(scores), _ptr(mask), batch, heads, seq_len, _f(scale), _ptr(probs), _ptr(dropped),
          _f(p_drop), C.c_uint(seed & 0xFFFFFFFF), _stream())
    return probs, dropped


def bert_softmax_bwd(probs, dprobs, batch, heads, seq_len, scale, p_drop=0.0, seed=0):
    ds = torch.empty(probs.shape, device=probs.device, dtype=torch.bfloat16)
    _call("ctclip_bert_softmax_bwd", _ptr(probs), _ptr(dprobs), batch, heads, seq_len, _f(scale), _ptr(ds), _f(p_drop),
          C.c_uint(seed & 0xFFFFFFFF), _stream())
    return ds


def gelu_fwd(h):
    _req(h, torch.bfloat16, "gelu_fwd.h")
    out = torch.empty_like(h)
    _call("ctclip_gelu_fwd", _ptr(h), _ptr(out), _ll(h.numel()), _stream())
    return out


def gelu_bwd(h, dy):
    dh = torch.empty_like(h)
    _call("ctclip_gelu_bwd", _ptr(h), _ptr(dy), _ptr(dh), _ll(h.numel()), _stream())
    return dh


def dropout_add(y, resid, p_drop, seed):
    """dropout(y) (+ resid); with resid=None this is also the dropout gradient (same seed -> same mask)"""
    _req(y, torch.float32, "dropout_add.y")
    out = torch.empty_like(y)
    _call("ctclip_dropout_add", _ptr(y), _ptr(resid), _ptr(out), _ll(y.numel()), _f(p_drop), C.c_uint(seed & 0xFFFFFFFF),
          _stream())
    return out


def colsum_bf16(x, out, dim=None, ld=None):
    _req(x, torch.bfloat16, "colsum_bf16.x")
    rows = x.shape[0]
    _call("ctclip_colsum_bf16", _ptr(x), _ll(rows), dim if dim is not None else x.shape[1],
          _ll(ld if ld is not None else x.stride(0)), _ptr(out), _stream())


def bert_attn_fwd(qkv, mask, B, H, L, D, p_drop=0.0, seed=0):
    """fused BertSelfAttention core (head dim 64): returns (context bf16 [B*L, D], lse fp32 [B*H, L])"""
    _req(qkv, torch.bfloat16, "bert_attn_fwd.qkv")
    assert qkv.is_contiguous() and qkv.shape == (B * L, 3 * D) and mask.dtype == torch.long and mask.is_contiguous()
    out = torch.empty((B * L, D), device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty((B * H, L), device=qkv.device, dtype=torch.float32)
    _call("ctclip_bert_attn_fwd", _ptr(qkv), _ptr(mask), B, H, L, D, _ptr(out), _ptr(lse), _f(p_drop),
          C.c_uint(seed & 0xFFFFFFFF), _stream())
    return out, lse


def bert_attn_bwd(qkv, mask, out, lse, dout, B, H, L, D, p_drop=0.0, seed=0):
    """returns the packed gradient dQKV bf16 [B*L, 3*D]"""
    _req(dout, torch.bfloat16, "bert_attn_bwd.dout")
    assert dout.is_contiguous() and out.is_contiguous() and lse.is_contiguous()
    dqkv = torch.empty_like(qkv)
    ws = torch.empty((B * H, L), device=qkv.device, dtype=torch.float32)
    _call("ctclip_bert_attn_bwd", _ptr(qkv), _ptr(mask), _ptr(out), _ptr(lse), _ptr(dout), B, H, L, D, _ptr(dqkv), _ptr(ws),
          _f(p_drop), C.c_uint(seed & 0xFFFFFFFF), _stream())
    return dqkv


def copy2d_batch(pairs, cache: dict):
    """pairs: list of (src, dst) 2-D bf16 tensors (same shape, last dim contiguous). One launch for all copies; the device
    descriptor table is cached in `cache` and reused while every address / shape is unchanged."""
    key = tuple((s.data_ptr(), d.data_ptr(), tuple(s.shape), s.stride(0), d.stride(0)) for s, d in pairs)
    if cache.get("copy2d_key") != key:
        rows = []
        for s, d in pairs:
            assert s.dtype == d.dtype == torch.bfloat16 and s.shape == d.shape and s.dim() == 2
            assert s.stride(1) == 1 and d.stride(1) == 1
            rows.append([s.data_ptr(), d.data_ptr(), s.shape[0], s.shape[1], s.stride(0), d.stride(0)])
        cache["copy2d_table"] = torch.tensor(rows, dtype=torch.int64).to(pairs[0][0].device)
        cache["copy2d_key"] = key
    _call("ctclip_copy2d_batch_bf16", _ptr(cache["copy2d_table"]), len(pairs), _stream())
