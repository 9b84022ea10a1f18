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
