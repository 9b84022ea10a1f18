"""GPU dev check: the CUDA CTCLIP (ctpa_clip_b200) against the CPU oracle on tiny/mid configs: forward + gradients."""
import sys, time
import torch
sys.path.insert(0, ".")
from transformers import BatchEncoding
from ctpa_clip_b200.ct_clip import CTViT, CTCLIP
from oracle import ctclip_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["tiny", "mid"]
allok = True

def build(cfg, sd, txt):
    vit = CTViT(dim=cfg["dim"], codebook_size=cfg["codebook_size"], image_size=cfg["image_size"], patch_size=cfg["patch_size"],
                temporal_patch_size=cfg["temporal_patch_size"], spatial_depth=cfg["spatial_depth"], temporal_depth=cfg["temporal_depth"],
                dim_head=cfg["dim_head"], heads=cfg["heads"])
    m = CTCLIP(image_encoder=vit, text_encoder=txt, dim_text=cfg["dim_text"], dim_image=cfg["dim_image"], dim_latent=cfg["dim_latent"])
    m.load_state_dict(sd, strict=False)
    return m.to(dev)

for name in names:
    cfg = O.CONFIGS[name]
    B = 3 if name != "production" else 2
    sd = O.init_state_dict(cfg, 0)
    txt_ref = O.make_text_encoder(cfg, 0)
    txt = O.make_text_encoder(cfg, 0)
    video, ids, mask = O.make_inputs(cfg, B, 0)
    # ---- oracle (fp32, on the GPU for speed at production size; same arithmetic as the CPU restatement)
    odev = "cuda" if name == "production" else "cpu"
    sdr = {k: v.clone().to(odev).requires_grad_(v.dtype.is_floating_point and v.numel() > 0 and "codebook" not in k and "beta" not in k)
           for k, v in sd.items()}
    txt_ref = txt_ref.to(odev)
    t0 = time.time()
    out = O.ctclip_forward(sdr, cfg, txt_ref, ids.to(odev), mask.to(odev), video.to(odev), training=True)
    out["loss"].backward()
    print(name, "oracle fwd+bwd s", time.time() - t0, "loss", float(out["loss"]))
    # ---- CUDA path
    m = build(cfg, sd, txt)
    m.text_autocast = (name == "production")
    m.train()
    if "--force" in sys.argv:
        m.visual_transformer.force_indices = out["indices"]
    text = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})
    loss = m(text, video.to(dev), return_loss=True)
    loss.backward()
    torch.cuda.synchronize()
    vit = m.visual_transformer
    idx = vit.last_indices.reshape(B, -1).cpu().long()
    agree = (idx == out["indices"].reshape(B, -1).cpu()).float().mean().item()
    dl = abs(float(loss) - float(out["loss"]))
    print(f"{name}: loss cuda {float(loss):.6f} oracle {float(out['loss']):.6f} |d|={dl:.2e}  vq index agreement {agree:.4f}")
    allok &= dl < 3e-2 * max(1.0, abs(float(out["loss"])))
    # latents (eval-style call on the same weights; EMA already mutated the CUDA codebook, so reload)
    m2 = build(cfg, sd, txt); m2.text_autocast = m.text_autocast; m2.eval()
    m2.visual_transformer.force_indices = m.visual_transformer.force_indices
    with torch.no_grad():
        tl, il, enc = m2(text, video.to(dev), return_latents=True)
    cos_t = torch.nn.functional.cosine_similarity(tl.cpu(), out["text_latents"].detach().cpu()).min().item()
    cos_i = torch.nn.functional.cosine_similarity(il.cpu(), out["image_latents"].detach().cpu()).min().item()
    print(f"{name}: min cos text {cos_t:.6f} image {cos_i:.6f}")
    allok &= cos_t > 0.99 and cos_i > 0.99
    # gradients
    worst = []
    for k, p in m.named_parameters():
        ref = None
        if k in sdr and sdr[k].grad is not None:
            ref = sdr[k].grad
        elif k.startswith("text_transformer."):
            rp = dict(txt_ref.named_parameters()).get(k[len("text_transformer."):])
            ref = rp.grad if rp is not None else None
        if ref is None:
            continue
        if p.grad is None:
            print("  MISSING GRAD", k); allok = False; continue
        g, r = p.grad.detach().float().cpu(), ref.detach().float().cpu()
        if r.norm() < 1e-6:
            continue
        rel = ((g - r).norm() / (r.norm() + 1e-12)).item()
        worst.append((rel, k, r.norm().item()))
    worst.sort(reverse=True)
    for rel, k, n in worst[:25]:
        print(f"  grad rel-err {rel:.3e}  |ref|={n:.3e}  {k}")
    med = sorted(w[0] for w in worst)[len(worst) // 2]
    print(f"{name}: {len(worst)} gradients compared, median rel-err {med:.3e}, max {worst[0][0]:.3e}")
    allok &= worst[0][0] < 0.2 and med < 5e-2
print("ALL_OK" if allok else "SOME_FAILED")
sys.exit(0 if allok else 1)
