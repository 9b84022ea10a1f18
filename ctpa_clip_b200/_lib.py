"""ctypes binding of libctclip_sm100.so (the C-ABI declared in include/ctclip_b200.h).

There is no CPU fallback: if the shared library is missing, `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libctclip_sm100.so"
_lib = None


class CtclipError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_longlong), ("a_mn_major", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_mn_major", C.c_int),
        ("C", C.c_void_p), ("ldc", C.c_longlong), ("c_is_f32", C.c_int),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_longlong),
        ("alpha", C.c_float),
        ("atomic", C.c_int),
        ("splits", C.c_int),
        ("top2_out", C.c_void_p),
        ("batch_h", C.c_int), ("batch_b", C.c_int),
        ("a_stride_h", C.c_longlong), ("a_stride_b", C.c_longlong),
        ("b_stride_h", C.c_longlong), ("b_stride_b", C.c_longlong),
        ("c_stride_h", C.c_longlong), ("c_stride_b", C.c_longlong),
        ("geglu_u", C.c_void_p), ("ld_u", C.c_longlong),
        ("geglu_h", C.c_void_p), ("ld_h", C.c_longlong),
    ]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise CtclipError(
                f"{LIB_PATH} not built: run `python -m ctpa_clip_b200.build` (there is no CPU/PyTorch fallback)")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.ctclip_version.restype = C.c_int
        _lib.ctclip_last_error.restype = C.c_int
        _lib.ctclip_last_error.argtypes = [C.c_char_p, C.c_size_t]
        _lib.ctclip_launch_count.restype = C.c_longlong
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().ctclip_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise CtclipError(f"{what} failed (rc={rc}): {last_error()}")


def launch_count() -> int:
    return int(lib().ctclip_launch_count())


class AttnDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("t", C.c_int), ("h", C.c_int), ("w", C.c_int), ("heads", C.c_int),
        ("dim_head", C.c_int), ("temporal", C.c_int),
        ("q", C.c_void_p), ("ldq", C.c_int),
        ("kv", C.c_void_p), ("ldkv", C.c_int),
        ("o", C.c_void_p), ("ldo", C.c_int),
        ("lse", C.c_void_p),
        ("q_scale", C.c_void_p), ("k_scale", C.c_void_p),
        ("bias_table", C.c_void_p), ("bias_rowmax", C.c_void_p),
        ("d_o", C.c_void_p),
        ("dq", C.c_void_p), ("dkv", C.c_void_p),
        ("dq_scale", C.c_void_p), ("dk_scale", C.c_void_p), ("dbias_table", C.c_void_p),
    ]


class PrepDesc(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p), ("out", C.c_void_p),
        ("batch", C.c_int), ("in_is_i16", C.c_int),
        ("slope", C.c_double), ("intercept", C.c_double),
        ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("stride_d", C.c_longlong), ("stride_h", C.c_longlong), ("stride_w", C.c_longlong),
        ("stride_batch", C.c_longlong),
        ("oD", C.c_int), ("oH", C.c_int), ("oW", C.c_int),
        ("tD", C.c_int), ("tH", C.c_int), ("tW", C.c_int),
        ("pad_value", C.c_float),
        ("lut_workspace", C.c_void_p),
        ("force_generic", C.c_int),
        ("pre_op", C.c_int), ("post_op", C.c_int),
    ]
