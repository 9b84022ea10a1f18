"""Worker of tests/test_gpu_peer_loss.py: one rank of the peer-memory latent exchange (csrc/symm.cu).
usage: python _peer_loss_worker.py RANK WORLD PORT B_LOCAL D STEPS
Ranks share cuda:0 when the box has one GPU (CUDA IPC works between processes on one device; the spin-waits are
time-sliced), otherwise rank r uses cuda:r. Rendezvous over gloo so that it also runs with two ranks on one device."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch
import torch.distributed as dist


def model_mode(rank, world):
    """tiny CT-CLIP, global batch world*2 split over the ranks: the data-parallel loss (fused exchange inside
    ClipLossFunction) must equal the single-process loss on the concatenated batch, and backward must run."""
    from transformers import BatchEncoding
    from ctpa_clip_b200 import ops
    from ctpa_clip_b200.ct_clip import CTCLIP, CTViT
    from ctpa_clip_b200.ct_clip.ct_clip import L2NormFunction
    from oracle import ctclip_oracle as O
    cfg = O.TINY
    b = 2
    vit = CTViT(dim=cfg["dim"], codebook_size=cfg["codebook_size"], image_size=cfg["image_size"], patch_size=cfg["patch_size"],
                temporal_patch_size=cfg["temporal_patch_size"], spatial_depth=cfg["spatial_depth"],
                temporal_depth=cfg["temporal_depth"], dim_head=cfg["dim_head"], heads=cfg["heads"])
    m = CTCLIP(image_encoder=vit, text_encoder=O.make_text_encoder(cfg, 0), dim_text=cfg["dim_text"],
               dim_image=cfg["dim_image"], dim_latent=cfg["dim_latent"])
    m.load_state_dict(O.init_state_dict(cfg, 0), strict=False)
    m.text_autocast = False
    m = m.cuda().eval()
    video, ids, mask = O.make_inputs(cfg, world * b, 0)
    video, ids, mask = video.cuda(), ids.cuda(), mask.cuda()
    with torch.no_grad():                                            # whole batch, no collective: latents -> single-device loss
        tl, il, _ = m(BatchEncoding({"input_ids": ids, "attention_mask": mask}), video, return_latents=True)
        tau = m.temperature.detach().reshape(1).float()
        want = ops.clip_loss(tl.float().contiguous(), il.float().contiguous(), tau, 0, world * b, want_grad=False)[0]
    sl = slice(rank * b, (rank + 1) * b)
    loss = m(BatchEncoding({"input_ids": ids[sl], "attention_mask": mask[sl]}), video[sl].contiguous(), return_loss=True)
    loss.backward()
    torch.cuda.synchronize()
    assert m.to_visual_latent.weight.grad is not None and torch.isfinite(m.to_visual_latent.weight.grad).all()
    # per-rank vs whole-batch GEMM shapes may pick different split-K orders -> 5e-3, not bit-exact
    assert abs(float(loss) - float(want)) < 5e-3, f"rank {rank}: data-parallel loss {float(loss)} vs single-process {float(want)}"
    return float(loss)


def main():
    rank, world, port, b, d, steps = (int(a) for a in sys.argv[1:7])
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(rank % ndev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ctpa_clip_b200 import _lib, ops, symm

    if steps == 0:                                                     # end-to-end mode through the CTCLIP module
        val = model_mode(rank, world)
        dist.barrier()
        symm.shutdown()
        dist.destroy_process_group()
        print(f"PEER_LOSS_OK rank {rank}/{world}: model loss {val:.6f} equals the single-process loss", flush=True)
        return
    ex = symm.LatentExchange(b, d)
    n0 = _lib.launch_count()
    dev = "cuda"
    for step in range(steps):
        g = torch.Generator().manual_seed(100 + step)                  # every rank knows the whole global batch
        T_all = torch.nn.functional.normalize(torch.randn(world * b, d, generator=g), dim=-1).to(dev)
        I_all = torch.nn.functional.normalize(torch.randn(world * b, d, generator=g), dim=-1).to(dev)
        tau = torch.tensor([0.3 + 0.2 * step], device=dev)
        sl = slice(rank * b, (rank + 1) * b)
        want = ops.clip_loss(T_all, I_all, tau, rank * b, b, want_grad=True)      # single-device kernel on the full batch
        got = ops.clip_loss_allgather(T_all[sl].contiguous(), I_all[sl].contiguous(), tau, rank, world, ex.table,
                                      ex.next_step())
        torch.cuda.synchronize()
        for name, a, w in zip(("loss", "dT", "dI", "dtau"), got, want):
            assert torch.isfinite(a).all(), f"rank {rank} step {step}: {name} not finite (a peer never arrived?)"
            if name == "dtau":      # per-row shares are added with fp32 atomics: the order (not the terms) may differ
                assert torch.allclose(a, w, rtol=1e-5, atol=1e-7), f"rank {rank} step {step}: dtau {a.item()} vs {w.item()}"
            else:
                assert torch.equal(a, w), f"rank {rank} step {step}: {name} differs, max |d| = {(a - w).abs().max().item():.3e}"
    launches = _lib.launch_count() - n0
    dist.barrier()
    ex.close()
    dist.destroy_process_group()
    print(f"PEER_LOSS_OK rank {rank}/{world} on cuda:{rank % ndev}, {steps} steps bit-exact, {launches} launches", flush=True)


if __name__ == "__main__":
    main()
