"""Shapes of the CT-CLIP configurations the hot path is built, tested and measured on.

PRODUCTION is CTPA_CLIP/ct_clip/pretrained_model.py:17-42 (CTViT dim 512, patch 20x20x10, 4+4 layers, 8x32 heads, codebook
8192; BERT-base text tower; 294912 -> 512 latent projection) on 480x480x240 volumes and 512-token reports — the
configuration BASELINE.json's metric is quoted on. TINY / MID are the small parity configurations (non-cubic and cubic
token grids). The CPU oracle keeps its own copy (oracle/ctclip_oracle.py); tests/test_abi_and_host.py asserts they agree.
"""
from __future__ import annotations

PRODUCTION = dict(
    dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10, spatial_depth=4,
    temporal_depth=4, dim_head=32, heads=8, frames=240, dim_text=768, dim_image=294912, dim_latent=512,
    text=dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
              intermediate_size=3072, max_position_embeddings=512), seq_len=512)
TINY = dict(  # token grid (t,h,w) = (5,4,4)
    dim=64, codebook_size=128, image_size=80, patch_size=20, temporal_patch_size=10, spatial_depth=2,
    temporal_depth=2, dim_head=32, heads=2, frames=50, dim_text=64, dim_image=4 * 4 * 64, dim_latent=32,
    text=dict(vocab_size=1000, hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
              max_position_embeddings=32), seq_len=16)
MID = dict(   # token grid (6,6,6), 4 heads x 32
    dim=128, codebook_size=512, image_size=120, patch_size=20, temporal_patch_size=10, spatial_depth=2,
    temporal_depth=2, dim_head=32, heads=4, frames=60, dim_text=64, dim_image=6 * 6 * 128, dim_latent=64,
    text=dict(vocab_size=1000, hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
              max_position_embeddings=32), seq_len=16)
CONFIGS = {"production": PRODUCTION, "tiny": TINY, "mid": MID}


def build_model(cfg: dict, device=None, seed: int | None = 0):
    """CTCLIP(CTViT, random-init HF BertModel) of `cfg` (CXR-BERT weights are not available offline)"""
    import torch
    from transformers import BertConfig, BertModel

    from .ct_clip import CTCLIP, CTViT
    if seed is not None:
        torch.manual_seed(seed)
    vit = CTViT(dim=cfg["dim"], codebook_size=cfg["codebook_size"], image_size=cfg["image_size"],
                patch_size=cfg["patch_size"], temporal_patch_size=cfg["temporal_patch_size"],
                spatial_depth=cfg["spatial_depth"], temporal_depth=cfg["temporal_depth"], dim_head=cfg["dim_head"],
                heads=cfg["heads"])
    txt = BertModel(BertConfig(**cfg["text"]))
    model = CTCLIP(image_encoder=vit, text_encoder=txt, dim_text=cfg["dim_text"], dim_image=cfg["dim_image"],
                   dim_latent=cfg["dim_latent"])
    return model.to(device) if device is not None else model


def synth_batch(cfg: dict, batch: int, seed: int):
    """seeded synthetic inputs: volumes U(-1,1) fp32 (b,1,f,h,w); token ids with the second half padding"""
    import torch
    g = torch.Generator().manual_seed(seed)
    video = torch.rand(batch, 1, cfg["frames"], cfg["image_size"], cfg["image_size"], generator=g) * 2 - 1
    L = cfg["seq_len"]
    ids = torch.randint(1, cfg["text"]["vocab_size"], (batch, L), generator=g)
    mask = torch.ones(batch, L, dtype=torch.long)
    ids[:, L // 2:] = 0
    mask[:, L // 2:] = 0
    return video, ids, mask
