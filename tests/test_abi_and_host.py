"""CPU: the C-ABI library loads and exports every symbol include/ctclip_b200.h declares; host-side logic."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "ctclip_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctclip_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ctpa_clip_b200 import _lib
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ctclip_b200.h but not exported"
    assert lib.ctclip_version() >= 100


def test_no_cpu_fallback_ops_refuse_cpu_tensors():
    from ctpa_clip_b200 import _lib, ops
    with pytest.raises(_lib.CtclipError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(_lib.CtclipError):
        ops.layernorm_fwd(torch.zeros(4, 8), torch.ones(8))


def test_struct_layouts_match_header():
    """ctypes mirrors have the same size as the C structs (compiled probe)"""
    import subprocess, tempfile
    from ctpa_clip_b200 import _lib
    src = '#include "ctclip_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(ctclip_gemm_desc), sizeof(ctclip_attn_desc), sizeof(ctclip_prep_desc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        p = Path(d) / "p.c"
        p.write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(p), "-o", str(Path(d) / "p")], check=True)
        out = subprocess.run([str(Path(d) / "p")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(_lib.GemmDesc), ctypes.sizeof(_lib.AttnDesc), ctypes.sizeof(_lib.PrepDesc)]


def test_state_dict_keys_match_reference_contract():
    """key names and shapes of the drop-in modules (tiny config)"""
    from ctpa_clip_b200.ct_clip import CTCLIP, CTViT
    from oracle import ctclip_oracle as O
    cfg = O.TINY
    vit = CTViT(dim=cfg["dim"], codebook_size=cfg["codebook_size"], image_size=cfg["image_size"], patch_size=cfg["patch_size"],
                temporal_patch_size=cfg["temporal_patch_size"], spatial_depth=cfg["spatial_depth"],
                temporal_depth=cfg["temporal_depth"], dim_head=cfg["dim_head"], heads=cfg["heads"])
    m = CTCLIP(image_encoder=vit, text_encoder=O.make_text_encoder(cfg), dim_text=cfg["dim_text"], dim_image=cfg["dim_image"],
               dim_latent=cfg["dim_latent"])
    sd = O.init_state_dict(cfg)
    mine = m.state_dict()
    for k, v in sd.items():
        assert k in mine and tuple(mine[k].shape) == tuple(v.shape), k
    for k in ("to_visual_latent_extra.weight", "to_text_latent_extra.weight", "visual_transformer.to_pixels.0.weight",
              "visual_transformer.to_patch_emb_first_frame.2.weight",
              "visual_transformer.enc_spatial_transformer.layers.0.1.null_kv",
              "visual_transformer.enc_temporal_transformer.norm_out.beta"):
        assert k in mine, k
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected


def test_unsupported_configs_refuse_loudly():
    from ctpa_clip_b200.ct_clip import CTCLIP, CTViT
    with pytest.raises(NotImplementedError):
        CTViT(dim=64, codebook_size=16, image_size=40, patch_size=20, temporal_patch_size=10, spatial_depth=1,
              temporal_depth=1, dim_head=64, heads=2)
    with pytest.raises(NotImplementedError):
        CTCLIP(image_encoder=None, text_encoder=None)


def test_unused_parameter_set_is_static():
    from ctpa_clip_b200.trainer import trainable_parameters
    from ctpa_clip_b200.ct_clip import CTCLIP, CTViT
    from oracle import ctclip_oracle as O
    cfg = O.TINY
    vit = CTViT(dim=cfg["dim"], codebook_size=cfg["codebook_size"], image_size=cfg["image_size"], patch_size=cfg["patch_size"],
                temporal_patch_size=cfg["temporal_patch_size"], spatial_depth=cfg["spatial_depth"],
                temporal_depth=cfg["temporal_depth"], dim_head=cfg["dim_head"], heads=cfg["heads"])
    m = CTCLIP(image_encoder=vit, text_encoder=O.make_text_encoder(cfg), dim_text=cfg["dim_text"], dim_image=cfg["dim_image"],
               dim_latent=cfg["dim_latent"])
    names = [n for n, _ in trainable_parameters(m)]
    fx = torch.load("tests/golden/ctclip_tiny.pt", weights_only=False)
    with_grad = set(fx["grads"].keys())            # parameters the REFERENCE's backward touches
    assert set(names) == with_grad


def test_workspace_bytes_host_arithmetic():
    """ctclip_workspace_bytes is pure host code: sizes equal the buffers ctpa_clip_b200.ops allocates"""
    from ctpa_clip_b200 import _lib
    lib = _lib.lib()
    lib.ctclip_workspace_bytes.restype = ctypes.c_longlong

    def ws(op, *dims):
        arr = (ctypes.c_longlong * max(1, len(dims)))(*dims)
        return lib.ctclip_workspace_bytes(op.encode(), arr, len(dims))

    assert ws("clip_loss", 64, 512) == (64 * 64 + 128 + 512) * 4
    assert ws("clip_loss_allgather", 8, 512, 8) == (64 * 64 + 128 + 512) * 4
    R = 47 * 47
    assert ws("cpb_table_fwd", 24, 24, 512) == (2 * R + 2 * R * 512) * 4
    assert ws("cpb_table_bwd", 24, 24, 512) == 2 * R * 512 * 4
    assert ws("bert_attn_bwd", 8, 12, 512) == 8 * 12 * 512 * 4
    assert ws("prep_resample") == 8192 * 4
    assert ws("sumsq") == 1024 * 4
    assert ws("no_such_op", 1) == -1
    assert "unknown op" in _lib.last_error()


def test_symmetric_buffer_layout_size():
    """ctclip_symm_latent_bytes: flag header + two parities of [T; I] for the whole global batch (host arithmetic only)"""
    from ctpa_clip_b200 import _lib
    lib = _lib.lib()
    lib.ctclip_symm_latent_bytes.restype = ctypes.c_size_t
    assert lib.ctclip_symm_latent_bytes(8, 512, 8) == (2 * 32 * 8 + 2 * 2 * 64 * 512) * 4
    assert lib.ctclip_symm_latent_bytes(8, 512, 33) == 0          # more ranks than the flag header holds
    assert lib.ctclip_symm_latent_bytes(0, 512, 2) == 0


def test_package_configs_equal_the_oracles():
    """bench.py's product arm takes its shapes from ctpa_clip_b200.configs (it imports nothing from oracle/)"""
    from ctpa_clip_b200 import configs
    from oracle import ctclip_oracle as O
    assert configs.CONFIGS == O.CONFIGS
    v, ids, mask = configs.synth_batch(configs.TINY, 2, 5)
    assert v.shape == (2, 1, 50, 80, 80) and ids.shape == (2, 16) and int(mask[:, 8:].sum()) == 0
    src = (ROOT / "bench.py").read_text()
    ours = src[src.index("def run_ours"):src.index('if __name__ == "__main__"')]
    assert not re.search(r"^\s*(import|from)\s+oracle", ours, flags=re.M), "product arm must not import oracle/"


def test_bias_table_pitch_is_conflict_free():
    """csrc/attention.cu tab_pitch(): the relative-position bias table rows sit at pitch S = gw + 32 * ceil((gw - 1) / 32) in
    shared memory. Lanes of a warp are 32 consecutive query positions p = qy * gw + qx reading word (qy + gh - 1) * S + qx +
    gw - 1 - B_j: for every grid and every start position the 32 words must fall into 32 different banks, every index must stay
    inside the padded table, and S must hold a whole table row."""
    def pitch(gw):
        return gw + 32 * ((gw - 1 + 31) // 32)
    for gh, gw in [(24, 24), (13, 11), (6, 6), (4, 4), (5, 5), (3, 4), (40, 33), (2, 64), (7, 1)]:
        S = pitch(gw)
        assert S >= 2 * gw - 1 and S % 32 == gw % 32
        words = (2 * gh - 2) * S + 2 * gw - 1
        n = gh * gw
        a = [((p // gw) + gh - 1) * S + (p % gw) + gw - 1 for p in range(n)]
        b = [(k // gw) * S + (k % gw) for k in range(n)]
        assert min(a) - max(b) == 0 and max(a) - min(b) == words - 1
        for start in range(0, max(1, n - 31)):
            lanes = a[start:start + 32]
            assert len({x % 32 for x in lanes}) == len(lanes), (gh, gw, start)


def test_pack12_round_trip_on_the_host():
    """data_prep/pack12.py: two voxels in three bytes; the inverse restores every value inside [-offset, 4095 - offset] and
    clamps the rest (which the HU window [-1000, 1000] of process_file would clip anyway)"""
    import numpy as np
    from ctpa_clip_b200.data_prep.pack12 import pack12, unpack12_numpy
    rng = np.random.default_rng(3)
    raw = rng.integers(-3000, 6000, size=(4, 8, 16), dtype=np.int16)
    raw.reshape(-1)[:6] = [-1024, 3071, -1025, 3072, 0, -1]
    for offset in (1024, 0, 2048):
        packed = pack12(raw, offset)
        assert packed.dtype.is_floating_point is False and packed.numel() == raw.size * 3 // 2
        back = unpack12_numpy(packed, offset).reshape(raw.shape)
        assert (back == np.clip(raw, -offset, 4095 - offset)).all()
