"""GPU mirror of the numeric core of CTPA_CLIP/data_prep/preprocess_train.py (== preprocess_test.py) and of the
resampling / crop / pad in ct_clip/data.py. File I/O (NIfTI, npz, CSV) stays with the caller.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops

TARGET_SPACING = (1.5, 0.75, 0.75)          # (z, x, y) — preprocess_train.py:92-97
TARGET_SHAPE = (240, 480, 480)              # (d, h, w) — data.py:153 target_shape (480,480,240) permuted


def resize_shape(shape, current_spacing, target_spacing):
    """new_shape[i] = int(shape[i] * current[i] / target[i]) in python float64 (preprocess_train.py:33-39)"""
    return [int(shape[i] * (current_spacing[i] / target_spacing[i])) for i in range(len(shape))]


def resize_array(array, current_spacing, target_spacing):
    """Same signature / return type as the reference resize_array (preprocess_train.py:31-42, ct_clip/data.py:15-40):
    (1, 1, D, H, W) float tensor -> (1, 1, D', H', W') np.ndarray, trilinear, align_corners=False — computed on the GPU."""
    assert array.dim() == 5 and array.shape[0] == 1 and array.shape[1] == 1
    new_shape = resize_shape(array.shape[2:], current_spacing, target_spacing)
    vol = array[0].to(device="cuda", dtype=torch.float32).contiguous()
    out = ops.prep_resample(vol, new_shape, layout="dhw")
    return out[None].cpu().numpy()


def preprocess_volumes(raw, slope, intercept, xy_spacing, z_spacing, target_spacing=TARGET_SPACING, target_shape=None,
                       out=None):
    """process_file arithmetic for a batch of same-shape raw scans (preprocess_train.py:99-109).
    raw: int16 CUDA tensor [b, H, W, N] in NIfTI array order. Returns fp32 [b, D', H', W'] (or target_shape with the
    data.py:155-190 centre-crop / pad(-1) applied)."""
    b, H, W, N = raw.shape
    new_shape = resize_shape((N, H, W), (z_spacing, xy_spacing, xy_spacing), target_spacing)
    return ops.prep_resample(raw, new_shape, hu=(slope, intercept), layout="hwn", target=target_shape, out=out)


def to_training_volume(vol_dhw, target_shape=TARGET_SHAPE, pad_value=-1.0):
    """centre crop / pad(-1) of an already normalised (b, D, H, W) fp32 volume to (b, 1, 240, 480, 480) (data.py:155-190)"""
    out = ops.prep_resample(vol_dhw.contiguous(), vol_dhw.shape[1:], layout="dhw", target=target_shape, pad_value=pad_value)
    return out[:, None]


def _as_f32_cuda(arr):
    t = torch.from_numpy(arr) if isinstance(arr, np.ndarray) else arr
    if t.dtype != torch.float32 or t.dim() != 3:
        raise ValueError("loader volumes are 3-D float32 arrays (the .npz files preprocess_train.py writes)")
    return t.to("cuda", non_blocking=True).contiguous()


def training_loader_volume(ct_scan, slope, intercept, xy_spacing, z_spacing, target_shape=TARGET_SHAPE):
    """CTReportDataset.npz_img_to_tensor (ct_clip/data.py:114-192) after the file / CSV reads, in ONE kernel:
    ct_scan is the float32 (H, W, N) array of the .npz; slope * x + intercept (float32), resize_array to spacing
    (1.5, 0.75, 0.75), clip [-1000, 1000] / 1000, centre crop / pad(-1) to (480, 480, 240), permute.
    Returns the (1, 240, 480, 480) CUDA tensor the Dataset yields, bit-identical to the reference's."""
    x = _as_f32_cuda(ct_scan)
    H, W, N = x.shape
    new_shape = resize_shape((N, H, W), (z_spacing, xy_spacing, xy_spacing), TARGET_SPACING)
    return ops.prep_resample(x[None], new_shape, hu=(slope, intercept), layout="hwn", target=target_shape, pad_value=-1.0,
                             pre_op="affine", post_op="clip_div")


def inference_loader_volume(img_data, target_shape=TARGET_SHAPE):
    """CTReportDatasetinfer.nii_img_to_tensor (ct_clip/data_inference.py:78-122) after np.load: img_data is the float32
    (D, H, W) array of the .npz; (clip(x * 1000, -1000, 200) + 400) / 600, centre crop / pad(-1), no resample.
    Returns (1, 240, 480, 480)."""
    x = _as_f32_cuda(img_data)
    return ops.prep_resample(x[None], x.shape, layout="dhw", target=target_shape, pad_value=-1.0, pre_op="infer_window")


def resample_to_target(image_tensor, target_depth=240, target_size=480):
    """The report generator's direct-to-target resample (ctpa_report/vqa_meditron.py:165-175, ct_scan_inference.py:50-55,
    data_utils.py:54-59): F.interpolate((C, D, H, W)[None], size=(240, 480, 480), mode='trilinear', align_corners=False)[0]
    on the GPU; a volume already at the target size is returned unchanged, like the reference."""
    if image_tensor.dim() == 3:
        image_tensor = image_tensor[None]
    C, D, H, W = image_tensor.shape
    tgt = (target_depth, target_size, target_size)
    x = image_tensor.to(device="cuda", dtype=torch.float32).contiguous()
    if (D, H, W) == tgt:
        return x
    return ops.prep_resample(x, tgt, layout="dhw")
