"""CPU: the C data_prep oracle against fixtures written by the reference's own resize_array / numpy lines."""
import hashlib

import numpy as np
import pytest

from oracle import resample_oracle as R

GOLD = np.load("tests/golden/resample.npz")


@pytest.mark.parametrize("pair", [(512, 480), (320, 240), (300, 240), (361, 240), (100, 333), (24, 24)])
def test_tap_tables_bit_exact(pair):
    n_in, n_out = pair
    wmat = GOLD[f"taps_{n_in}_{n_out}"]
    i0, i1, w0, w1 = R.taps(n_in, n_out)
    mine = np.zeros_like(wmat)
    for o in range(n_out):
        mine[o, i0[o]] += w0[o]
        mine[o, i1[o]] += w1[o]
    assert (mine.view(np.int32) == wmat.view(np.int32)).all()


@pytest.mark.parametrize("case", range(4))
def test_trilinear_values_bit_exact(case):
    x, cur, want = GOLD[f"case{case}_in"], GOLD[f"case{case}_cur"], GOLD[f"case{case}_out"]
    shape = R.resize_shape(x.shape, tuple(cur), (1.5, 0.75, 0.75))
    got = R.trilinear(np.ascontiguousarray(x), shape)
    assert got.shape == want.shape
    assert (got.view(np.int32) == want.view(np.int32)).all()


def test_production_slab_hash():
    rng = np.random.default_rng(2)
    for shape in [(20, 32, 32), (17, 25, 23), (12, 24, 24), (9, 40, 30)]:  # replay the generator's stream
        rng.random(shape, dtype=np.float32)
    x = (rng.random((16, 512, 512), dtype=np.float32) * 2 - 1)
    assert hashlib.sha256(x.tobytes()).digest() == GOLD["slab_in_sha256"].tobytes()
    got = R.trilinear(x, R.resize_shape(x.shape, (1.125, 0.703125, 0.703125), (1.5, 0.75, 0.75)))
    assert tuple(got.shape) == tuple(GOLD["slab_shape"]) == (12, 480, 480)
    assert hashlib.sha256(got.tobytes()).digest() == GOLD["slab_sha256"].tobytes()


@pytest.mark.parametrize("j", range(3))
def test_hu_normalise_bit_exact(j):
    slope, intercept = GOLD[f"hu{j}_params"]
    got = R.hu_normalise(np.ascontiguousarray(GOLD["hu_raw"]), float(slope), float(intercept))
    assert (got.view(np.int32) == GOLD[f"hu{j}_out"].view(np.int32)).all()


def test_resize_shape_truncation():
    assert R.resize_shape((320, 512, 512), (1.125, 0.703125, 0.703125), (1.5, 0.75, 0.75)) == [240, 480, 480]
    assert R.resize_shape((300, 512, 512), (1.2, 0.75, 0.75), (1.5, 0.75, 0.75))[0] == 239  # fp truncation


def test_crop_pad_matches_data_py_arithmetic():
    """numpy transcription of ct_clip/data.py:155-190 on a small volume"""
    import torch
    rng = np.random.default_rng(0)
    for shape, target in [((10, 30, 21), (8, 24, 24)), ((5, 20, 30), (8, 24, 24)), ((8, 24, 24), (8, 24, 24))]:
        vol = rng.random(shape, dtype=np.float32)           # (D, H, W)
        tensor = torch.tensor(vol).permute(1, 2, 0)         # data.py works on (H, W, D)
        dh, dw, dd = target[1], target[2], target[0]
        h, w, d = tensor.shape
        h_start, h_end = max((h - dh) // 2, 0), min((h - dh) // 2 + dh, h)
        w_start, w_end = max((w - dw) // 2, 0), min((w - dw) // 2 + dw, w)
        d_start, d_end = max((d - dd) // 2, 0), min((d - dd) // 2 + dd, d)
        tensor = tensor[h_start:h_end, w_start:w_end, d_start:d_end]
        pads = []
        for want, have in ((dd, tensor.size(2)), (dw, tensor.size(1)), (dh, tensor.size(0))):
            pads += [(want - have) // 2, want - have - (want - have) // 2]
        tensor = torch.nn.functional.pad(tensor, tuple(pads), value=-1).permute(2, 0, 1)
        got = R.crop_pad(np.ascontiguousarray(vol), target, -1.0)
        assert np.array_equal(got, tensor.numpy())


def test_fp32_newton_division_equals_numpy_double_path_for_all_integer_hu():
    """prep_hwn_i16_kernel<ARITH>: for slope 1 and an integer intercept, c = clip(raw + intercept) is an integer in
    [-1000, 1000]; the kernel computes q0 = c * 0.001f, q1 = fma(fma(-q0, 1000, c), 0.001f, q0) in fp32. This must equal
    numpy's float32(float64(c) / 1000.0) (preprocess_train.py:103-105) for every one of the 2001 values (exact rational
    emulation of the fp32 roundings)."""
    import math
    from fractions import Fraction as F

    def rnd32(x):
        if x == 0:
            return F(0)
        s = -1 if x < 0 else 1
        x = abs(x)
        e = math.floor(math.log2(float(x))) - 23
        while x / F(2) ** e >= 2 ** 24:
            e += 1
        while x / F(2) ** e < 2 ** 23:
            e -= 1
        y = x / F(2) ** e
        m = math.floor(y)
        fr = y - m
        if fr > F(1, 2) or (fr == F(1, 2) and m % 2 == 1):
            m += 1
        return s * m * F(2) ** e

    r = rnd32(F(1, 1000))
    assert float(r) == float(np.float32(0.001))
    for c in range(-1000, 1001):
        want = np.float32(np.float64(c) / 1000.0)
        q0 = rnd32(F(c) * r)
        q1 = rnd32(rnd32(-q0 * 1000 + c) * r + q0)
        assert float(q1) == float(want), c


# ---- DataLoader conversions: oracle restatement pinned against the reference's own loaders ------------
def _loader_golden():
    return np.load("tests/golden/loaders.npz")


def _check_volume(out, g, key):
    import hashlib
    assert out.shape == (1, 240, 480, 480) and out.dtype == np.float32
    assert hashlib.sha256(out.tobytes()).digest() == g[key + "_sha256"].tobytes()
    if key + "_win" in g:
        lo, hi = g[key + "_box"][:3], g[key + "_box"][3:]
        assert np.array_equal(out[0, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]], g[key + "_win"])
        assert out[0, 0, 0, 0] == -1.0                           # pad value, data.py:186 / data_inference.py:115


@pytest.mark.parametrize("i", [0, 1])
def test_inference_loader_oracle_matches_reference(i):
    """golden = CTReportDatasetinfer.nii_img_to_tensor run verbatim (tools/make_golden.py loaders); case 1 crops depth"""
    g = _loader_golden()
    arr = g[f"infer{i}_q"].astype(np.float32) / np.float32(1024)
    _check_volume(R.inference_loader_volume(arr), g, f"infer{i}")


@pytest.mark.parametrize("i", [0, 1, 2])
def test_training_loader_oracle_matches_reference(i):
    """golden = CTReportDataset.npz_img_to_tensor run verbatim; case 2 has a slope / intercept that float32 cannot
    represent (pins the float32 weak-scalar arithmetic of data.py:138) and up-samples depth 2x"""
    g = _loader_golden()
    s, ic, xy, z = (float(v) for v in g[f"train{i}_params"])
    _check_volume(R.training_loader_volume(g[f"train{i}_in"], s, ic, xy, z), g, f"train{i}")
