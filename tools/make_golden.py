"""Writes tests/golden/*.pt / *.npz by running the UNMODIFIED reference (/root/reference, authoring container only)
through oracle/ref_shim.py on the seeded synthetic weights / inputs of oracle/ctclip_oracle.py.

    python tools/make_golden.py [tiny mid mid4 production resample loaders]

The fixtures pin (i) the oracle restatement against the real reference modules, (ii) the data_prep index/weight rule and
value arithmetic against the reference's own resize_array. They travel to the GPU box; /root/reference does not.
"""
import hashlib
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ctclip_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = ROOT / "tests" / "golden"
GOLD.mkdir(parents=True, exist_ok=True)


def probe_vector(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def model_fixture(name):
    """name: tiny | mid | production, or mid4 = the mid configuration on a batch of 4 (the concatenated global batch of the
    2- and 4-rank data-parallel parity test, tests/test_gpu_multi.py)"""
    from transformers import BatchEncoding
    tag = name
    batch = {"production": 2, "mid4": 4}.get(name, 3)
    name = "mid" if name == "mid4" else name
    cfg = O.CONFIGS[name]
    sd = O.init_state_dict(cfg, 0)
    txt = O.make_text_encoder(cfg, 0)
    model = ref_shim.build_reference_model(cfg, txt)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    video, ids, mask = O.make_inputs(cfg, batch, 0)
    text = BatchEncoding({"input_ids": ids, "attention_mask": mask})
    fx = {"config": name, "batch": batch, "seed": 0}
    model.eval()
    t0 = time.time()
    with torch.no_grad():
        fx["loss_eval"] = ref_shim.quiet(model, text, video, device="cpu", return_loss=True).clone()
        tl, il, enc = ref_shim.quiet(model, text, video, device="cpu", return_latents=True)
        fx["sim_eval"] = ref_shim.quiet(model, text, video, device="cpu").clone()
        vit = model.visual_transformer
        pre = vit.encode(vit.to_patch_emb(video))
        idx = vit(video, return_only_codebook_ids=True)
    print(name, "reference eval forward x4:", round(time.time() - t0, 1), "s")
    fx["text_latents"], fx["image_latents"] = tl.clone(), il.clone()
    fx["indices"] = idx.reshape(batch, -1).to(torch.int32)
    b, t, h, w, d = pre.shape
    pre = pre.reshape(batch, -1, d)
    fx["pre_vq_full"] = pre.clone() if tag in ("tiny", "mid") else None
    fx["pre_vq_head"] = pre[:, :32].clone()
    fx["pre_vq_rowsum"] = pre.sum(dim=-1)
    fx["enc_checksum"] = enc.double().sum().item()
    # train mode: loss + gradient summaries (+ EMA-updated codebook)
    model.train()
    txt.eval()
    model.zero_grad()
    t0 = time.time()
    loss = ref_shim.quiet(model, text, video, device="cpu", return_loss=True)
    loss.backward()
    print(name, "reference train fwd+bwd:", round(time.time() - t0, 1), "s")
    fx["loss_train"] = loss.detach().clone()
    grads = {}
    for i, (k, p) in enumerate(model.named_parameters()):
        if p.grad is None or p.numel() == 0:
            continue
        g = p.grad.detach()
        grads[k] = dict(norm=g.norm().item(), proj=(g * probe_vector(g.shape, 77 + i)).sum().item(), probe_seed=77 + i,
                        full=g.clone() if (g.numel() <= 4096 or name == "tiny") else None)
    fx["grads"] = grads
    cb = model.visual_transformer.vq._codebook
    fx["ema_cluster_size"] = cb.cluster_size.clone()
    fx["ema_embed_rowsum"] = cb.embed[0].sum(dim=-1).clone()
    fx["ema_embed_head"] = cb.embed[0, :16].clone()
    torch.save(fx, GOLD / f"ctclip_{tag}.pt")
    print(tag, "->", GOLD / f"ctclip_{tag}.pt", (GOLD / f"ctclip_{tag}.pt").stat().st_size // 1024, "KiB")


def resample_fixture():
    sys.path.insert(0, str(ref_shim.REFERENCE_ROOT))
    import torch.nn.functional as F
    from ct_clip.data import resize_array  # the reference's own function (ct_clip/data.py:15-40)
    out = {}
    # 1-D tap tables by one-hot probing of F.interpolate
    for n_in, n_out in [(512, 480), (320, 240), (300, 240), (361, 240), (100, 333), (24, 24)]:
        eye = torch.eye(n_in).reshape(n_in, 1, n_in, 1, 1)
        wmat = F.interpolate(eye, size=(n_out, 1, 1), mode="trilinear", align_corners=False)[:, 0, :, 0, 0].t().contiguous()
        out[f"taps_{n_in}_{n_out}"] = wmat.numpy()
    rng = np.random.default_rng(2)
    cases = [((20, 32, 32), (1.125, 0.703125, 0.703125)), ((17, 25, 23), (2.0, 0.6, 0.9)), ((12, 24, 24), (1.5, 0.75, 0.75)),
             ((9, 40, 30), (0.8, 1.3, 0.5))]
    for i, (shape, cur) in enumerate(cases):
        x = (rng.random(shape, dtype=np.float32) * 2 - 1)
        y = resize_array(torch.tensor(x)[None, None], cur, (1.5, 0.75, 0.75))[0][0]
        out[f"case{i}_in"], out[f"case{i}_cur"], out[f"case{i}_out"] = x, np.array(cur), y
    # one production-sized slab: hash only
    x = (rng.random((16, 512, 512), dtype=np.float32) * 2 - 1)
    y = resize_array(torch.tensor(x)[None, None], (1.125, 0.703125, 0.703125), (1.5, 0.75, 0.75))[0][0]
    out["slab_seed"] = np.array([2])
    out["slab_shape"] = np.array(y.shape)
    out["slab_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(y).tobytes()).digest(), dtype=np.uint8)
    out["slab_in_sha256"] = np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8)
    # HU normalisation (numpy lines of preprocess_train.py:99-104 executed verbatim)
    raw = rng.integers(-1024, 3071, size=(24, 24, 20), dtype=np.int16)
    for j, (slope, intercept) in enumerate([(1.0, 0.0), (1.0, -1024.0), (0.75, 13.5)]):
        img_data = raw.astype(np.float64)  # nib get_fdata() returns float64
        img_data = slope * img_data + intercept
        img_data = np.clip(img_data, -1000, 1000)
        img_data = ((img_data / 1000)).astype(np.float32)
        img_data = img_data.transpose(2, 0, 1)
        out[f"hu{j}_params"] = np.array([slope, intercept])
        out[f"hu{j}_out"] = np.ascontiguousarray(img_data)
    out["hu_raw"] = raw
    np.savez_compressed(GOLD / "resample.npz", **out)
    print("resample ->", (GOLD / "resample.npz").stat().st_size // 1024, "KiB")


def loaders_fixture():
    """the reference's two DataLoader conversions run verbatim on small synthetic .npz arrays (outputs are the fixed
    (1,240,480,480) training volume: stored as sha256 + the non-padding window)"""
    import os, tempfile, types
    import pandas as pd
    sys.path.insert(0, str(ref_shim.REFERENCE_ROOT))
    import ct_clip.data as rdata
    import ct_clip.data_inference as rinf
    rng = np.random.default_rng(7)
    out = {}

    def window_of(t):  # (1,240,480,480) -> bounding box of the values that are not the -1 padding
        a = t[0].numpy()
        nz = np.argwhere(a != -1.0)
        lo, hi = nz.min(0), nz.max(0) + 1
        return np.concatenate([lo, hi]), np.ascontiguousarray(a[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]])

    with tempfile.TemporaryDirectory() as td:
        # ---- inference loader (data_inference.py:78-122)
        for i, shape in enumerate([(20, 36, 30), (484, 4, 244)]):
            q = rng.integers(-1126, 512, size=shape, dtype=np.int16)     # stored as int16: arr = q / 1024 exactly
            arr = q.astype(np.float32) / np.float32(1024)
            path = os.path.join(td, f"inf{i}.npz")
            np.savez(path, arr)
            t = rinf.CTReportDatasetinfer.nii_img_to_tensor(types.SimpleNamespace(), path, None)
            box, win = window_of(t)
            out[f"infer{i}_q"], out[f"infer{i}_box"] = q, box
            if win.size < 100000:                                        # the crop case is pinned by its sha256 only
                out[f"infer{i}_win"] = win
            out[f"infer{i}_sha256"] = np.frombuffer(hashlib.sha256(t.numpy().tobytes()).digest(), dtype=np.uint8)
        # ---- training loader (data.py:114-192): metadata CSV row patched in
        cases = [((24, 30, 18), 1.0, 0.0, 0.9, 2.0), ((40, 26, 33), 1.0, -24.0, 0.6, 1.2), ((16, 20, 12), 1.0007, -10.3, 1.1, 3.0)]
        for i, (shape, slope, intercept, xy, z) in enumerate(cases):
            arr = (rng.random(shape, dtype=np.float32) * 2600 - 1300).astype(np.float32)
            path = os.path.join(td, f"train{i}.npz")
            np.savez(path, arr)
            df = pd.DataFrame({"VolumeName": [f"train{i}.nii"], "RescaleSlope": [slope], "RescaleIntercept": [intercept],
                               "XYSpacing": [f"[{xy}, {xy}]"], "ZSpacing": [z]})
            real = pd.read_csv
            pd.read_csv = lambda *a, **k: df
            try:
                t = rdata.CTReportDataset.npz_img_to_tensor(types.SimpleNamespace(split="train"), path, None)
            finally:
                pd.read_csv = real
            box, win = window_of(t)
            out[f"train{i}_in"], out[f"train{i}_params"] = arr, np.array([slope, intercept, xy, z])
            out[f"train{i}_box"], out[f"train{i}_win"] = box, win
            out[f"train{i}_sha256"] = np.frombuffer(hashlib.sha256(t.numpy().tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(GOLD / "loaders.npz", **out)
    print("loaders ->", (GOLD / "loaders.npz").stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    what = sys.argv[1:] or ["tiny", "mid", "production", "resample", "loaders"]
    assert ref_shim.available(), "needs /root/reference"
    for w in what:
        if w == "resample":
            resample_fixture()
        elif w == "loaders":
            loaders_fixture()
        else:
            model_fixture(w)
