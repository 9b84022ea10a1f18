"""helpers shared by the GPU tests and their multi-process workers (tests/ is on sys.path in both cases)"""


def build(cfg, sd, txt):
    """the drop-in CTCLIP(CTViT, injected text encoder) of `cfg` with the reference-keyed state_dict `sd` loaded, on cuda"""
    from ctpa_clip_b200.ct_clip import CTCLIP, CTViT
    vit = CTViT(dim=cfg["dim"], codebook_size=cfg["codebook_size"], image_size=cfg["image_size"], patch_size=cfg["patch_size"],
                temporal_patch_size=cfg["temporal_patch_size"], spatial_depth=cfg["spatial_depth"],
                temporal_depth=cfg["temporal_depth"], dim_head=cfg["dim_head"], heads=cfg["heads"])
    m = CTCLIP(image_encoder=vit, text_encoder=txt, dim_text=cfg["dim_text"], dim_image=cfg["dim_image"],
               dim_latent=cfg["dim_latent"])
    m.load_state_dict(sd, strict=False)
    m.text_autocast = False
    return m.cuda()


def text_of(ids, mask):
    from transformers import BatchEncoding
    return BatchEncoding({"input_ids": ids.cuda(), "attention_mask": mask.cuda()})
