"""Tensor-level wrappers over the C-ABI: take torch CUDA tensors, pass raw pointers + the current stream.

PyTorch is only the allocator / stream provider here; all arithmetic happens in libctclip_sm100.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream() -> C.c_void_p:
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
         alpha: float = 1.0, accumulate: bool = False, splits: int = 1) -> torch.Tensor:
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
    _lib.check(_lib.lib().ctclip_gemm_bf16(C.byref(d), _stream()), "ctclip_gemm_bf16")
    return out


def _call(name: str, *args):
    fn = getattr(_lib.lib(), name)
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
    _req(dy, torch.float32, "layernorm_bwd.dy")
    _req(x, torch.float32, "layernorm_bwd.x")
    rows, dim = x.shape
    assert dy.is_contiguous() and x.is_contiguous()
    dx = out if out is not None else torch.empty_like(x)
    dxb = torch.empty((rows, dim), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    _call("ctclip_layernorm_bwd", _ptr(dy), _ptr(x), _ll(rows), dim, _ptr(gamma), _f(eps), _ptr(add_in), _ptr(dx),
          _ptr(dxb), _ptr(dgamma), _ptr(dbeta), _stream())
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


def cast_bf16(x):
    _req(x, torch.float32, "cast_bf16.x")
    assert x.is_contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
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
    _call("ctclip_attn_fwd", C.byref(d), _stream())
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
    _call("ctclip_attn_bwd", C.byref(d), _stream())
    return dq, dkv
