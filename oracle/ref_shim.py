"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference from /root/reference (authoring container only).

Used solely by tools/make_golden.py to pin oracle/ctclip_oracle.py against the real reference modules and to
write the fixtures under tests/golden/. /root/reference does not exist on the GPU box, so nothing in tests/,
smoke() or bench.py may import this file at run time.

Recipe:
  1. register oracle/vq_restatement.py as the absent third-party `vector_quantize_pytorch` module (ctvit.py:17)
  2. put /root/reference/CTPA_CLIP on sys.path and import ct_clip.{ct_clip,ctvit,attention} unmodified
  3. neutralise the tokenizer download in CTCLIP.__init__ (ct_clip.py:585)
  4. on a GPU-less host rebind the module-global `torch` in ct_clip.attention / ct_clip.ctvit to a proxy whose
     .device(...) returns CPU (the reference hard-codes torch.device('cuda'): attention.py:135,260, ctvit.py:316,398)
"""
from __future__ import annotations

import contextlib
import io
import sys
import types
from pathlib import Path

import torch

REFERENCE_ROOT = Path("/root/reference/CTPA_CLIP")


def available() -> bool:
    return (REFERENCE_ROOT / "ct_clip" / "ct_clip.py").exists()


class _TorchCpuProxy:
    def __getattr__(self, name):
        return getattr(torch, name)

    def device(self, *a, **k):
        return torch.device("cpu")


_loaded = None


def load():
    """returns (ct_clip module, ctvit module, attention module) of the reference"""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?) — use the committed fixtures instead")
    here = Path(__file__).resolve().parent
    if str(here.parent) not in sys.path:
        sys.path.insert(0, str(here.parent))
    from oracle import vq_restatement

    mod = types.ModuleType("vector_quantize_pytorch")
    mod.VectorQuantize = vq_restatement.VectorQuantize
    sys.modules["vector_quantize_pytorch"] = mod
    sys.path.insert(0, str(REFERENCE_ROOT))
    import ct_clip.attention as att
    import ct_clip.ct_clip as cc
    import ct_clip.ctvit as cv

    class _Tok:
        @classmethod
        def from_pretrained(cls, *a, **k):
            return cls()

    cc.BertTokenizer = _Tok
    if not torch.cuda.is_available():
        att.torch = cv.torch = _TorchCpuProxy()
    _loaded = (cc, cv, att)
    return _loaded


def build_reference_model(cfg: dict, text_encoder: torch.nn.Module):
    """CTCLIP(image_encoder=CTViT(...), text_encoder=...) exactly as pretrained_model.py:17-42 wires it."""
    cc, cv, _ = load()
    enc = cv.CTViT(dim=cfg["dim"], codebook_size=cfg["codebook_size"], image_size=cfg["image_size"],
                   patch_size=cfg["patch_size"], temporal_patch_size=cfg["temporal_patch_size"],
                   spatial_depth=cfg["spatial_depth"], temporal_depth=cfg["temporal_depth"],
                   dim_head=cfg["dim_head"], heads=cfg["heads"], use_vgg_and_gan=False)
    model = cc.CTCLIP(image_encoder=enc, text_encoder=text_encoder, dim_text=cfg["dim_text"],
                      dim_image=cfg["dim_image"], dim_latent=cfg["dim_latent"], extra_latent_projection=False,
                      use_mlm=False, downsample_image_embeds=False, use_all_token_embeds=False)
    return model


def quiet(fn, *a, **k):
    """the reference prints 'test all pooling' on every forward (ct_clip.py:737)"""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
