"""TEST INFRASTRUCTURE — numpy/ctypes front end of oracle/resample_oracle.c (data_prep CPU oracle).

resize_shape follows CTPA_CLIP/data_prep/preprocess_train.py:33-39 (python float64 arithmetic, int() truncation).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .build_oracle import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    return _lib


def resize_shape(shape, current_spacing, target_spacing):
    return [int(shape[i] * (current_spacing[i] / target_spacing[i])) for i in range(len(shape))]


def hu_normalise(raw_hwn: np.ndarray, slope: float, intercept: float) -> np.ndarray:
    assert raw_hwn.dtype == np.int16 and raw_hwn.flags.c_contiguous
    H, W, N = raw_hwn.shape
    out = np.empty((N, H, W), np.float32)
    lib().ctclip_oracle_hu_normalise_i16(raw_hwn.ctypes.data_as(C.c_void_p), H, W, N, C.c_double(slope),
                                         C.c_double(intercept), out.ctypes.data_as(C.c_void_p))
    return out


def trilinear(vol_dhw: np.ndarray, out_shape) -> np.ndarray:
    assert vol_dhw.dtype == np.float32 and vol_dhw.flags.c_contiguous
    D, H, W = vol_dhw.shape
    oD, oH, oW = (int(s) for s in out_shape)
    out = np.empty((oD, oH, oW), np.float32)
    lib().ctclip_oracle_trilinear(vol_dhw.ctypes.data_as(C.c_void_p), D, H, W, out.ctypes.data_as(C.c_void_p), oD, oH,
                                  oW)
    return out


def taps(n_in: int, n_out: int):
    """(i0, i1, w0, w1) arrays of the 2-tap rule along one axis"""
    i0 = np.empty(n_out, np.int32); i1 = np.empty(n_out, np.int32)
    w0 = np.empty(n_out, np.float32); w1 = np.empty(n_out, np.float32)
    a, b, c, d = C.c_int(), C.c_int(), C.c_float(), C.c_float()
    for o in range(n_out):
        lib().ctclip_oracle_taps(n_in, n_out, o, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        i0[o], i1[o], w0[o], w1[o] = a.value, b.value, c.value, d.value
    return i0, i1, w0, w1


def crop_pad(vol_dhw: np.ndarray, target=(240, 480, 480), pad_value: float = -1.0) -> np.ndarray:
    assert vol_dhw.dtype == np.float32 and vol_dhw.flags.c_contiguous
    D, H, W = vol_dhw.shape
    out = np.empty(tuple(target), np.float32)
    lib().ctclip_oracle_crop_pad(vol_dhw.ctypes.data_as(C.c_void_p), D, H, W, out.ctypes.data_as(C.c_void_p),
                                 int(target[0]), int(target[1]), int(target[2]), C.c_float(pad_value))
    return out


def preprocess_volume(raw_hwn: np.ndarray, slope: float, intercept: float, xy_spacing: float, z_spacing: float,
                      target=(1.5, 0.75, 0.75)) -> np.ndarray:
    """process_file arithmetic, preprocess_train.py:92-109: normalise, then resize_array to the spacing-derived shape"""
    vol = hu_normalise(raw_hwn, slope, intercept)
    new_shape = resize_shape(vol.shape, (z_spacing, xy_spacing, xy_spacing), target)
    return trilinear(vol, new_shape)


def _unary(fn, x: np.ndarray, *args) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    getattr(lib(), fn)(x.ctypes.data_as(C.c_void_p), C.c_size_t(x.size), *args, out.ctypes.data_as(C.c_void_p))
    return out


def training_loader_volume(arr: np.ndarray, slope: float, intercept: float, xy_spacing: float, z_spacing: float,
                           target=(240, 480, 480)) -> np.ndarray:
    """CTReportDataset.npz_img_to_tensor, CTPA_CLIP/ct_clip/data.py:138-192, on the float32 array of an .npz:
    slope*x+intercept (float32) -> transpose(2,0,1) -> resize_array to spacing (1.5, 0.75, 0.75) -> clip[-1000,1000]/1000 ->
    centre crop / pad(-1) to (480,480,240) in (h,w,d) order -> permute to (1, 240, 480, 480)."""
    assert arr.dtype == np.float32 and arr.ndim == 3
    img = _unary("ctclip_oracle_affine_f32", arr, C.c_float(slope), C.c_float(intercept))
    img = np.ascontiguousarray(img.transpose(2, 0, 1))                                   # data.py:139
    new_shape = resize_shape(img.shape, (z_spacing, xy_spacing, xy_spacing), (1.5, 0.75, 0.75))
    img = trilinear(img, new_shape)                                                      # data.py:142-146
    img = _unary("ctclip_oracle_clip_div_f32", img)                                      # data.py:150-152
    return crop_pad(img, target, -1.0)[None]                                             # data.py:155-190


def inference_loader_volume(arr: np.ndarray, target=(240, 480, 480)) -> np.ndarray:
    """CTReportDatasetinfer.nii_img_to_tensor, CTPA_CLIP/ct_clip/data_inference.py:78-122: x*1000 -> clip[-1000,200] ->
    (x+400)/600 (float32) -> centre crop / pad(-1) to (480,480,240) -> (1, 240, 480, 480). No resample."""
    assert arr.dtype == np.float32 and arr.ndim == 3
    img = _unary("ctclip_oracle_window_infer", arr)
    return crop_pad(img, target, -1.0)[None]
