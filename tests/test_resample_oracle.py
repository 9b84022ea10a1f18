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
