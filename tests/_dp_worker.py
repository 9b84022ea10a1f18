"""Worker of tests/test_gpu_multi.py: one data-parallel rank on its OWN GPU (cuda:RANK), NCCL over NVLink.
usage: python _dp_worker.py RANK WORLD PORT MODE
MODE "train": the mid CT-CLIP configuration, global batch 4 split over the ranks, through CTClipTrainStep — the global-batch
InfoNCE (peer-memory latent exchange or NCCL all-gather, CTCLIP_LATENT_EXCHANGE), gradient reduction (buckets + factor gather
of the to_visual_latent gradient) and the all-reduced VQ EMA statistics — against the UNMODIFIED reference's single-process
loss / per-parameter gradients / EMA buffers on the concatenated batch (tests/golden/ctclip_mid4.pt, SURVEY §8(e) oracle).
MODE "kernel": 5 consecutive steps of ctclip_clip_loss_allgather, bit-exact against the single-device kernel."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch
import torch.distributed as dist


def probe(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def train_mode(rank, world):
    from transformers import BatchEncoding
    from ctpa_clip_b200 import symm
    from ctpa_clip_b200.trainer import CTClipTrainStep
    from oracle import ctclip_oracle as O
    from _common import build
    fx = torch.load("tests/golden/ctclip_mid4.pt", weights_only=False)
    cfg = O.MID
    B = fx["batch"]
    b = B // world
    # rank r > 0 starts from DIFFERENT weights on purpose: the trainer's rank-0 broadcast must make the replicas identical
    sd = O.init_state_dict(cfg, 0 if rank == 0 else 11)
    m = build(cfg, sd, O.make_text_encoder(cfg, 0 if rank == 0 else 11)).train()
    m.text_transformer.eval()                                      # fixture: BERT dropout off
    video, ids, mask = O.make_inputs(cfg, B, 0)
    sl = slice(rank * b, (rank + 1) * b)
    m.visual_transformer.force_indices = fx["indices"][sl]
    tr = CTClipTrainStep(m)
    loss = tr.forward_backward(BatchEncoding({"input_ids": ids[sl].cuda(), "attention_mask": mask[sl].cuda()}),
                               video[sl].contiguous().cuda())
    tr.reduce_gradients()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(fx["loss_train"])) < 2e-3, (float(loss), float(fx["loss_train"]))
    params = dict(m.named_parameters())
    worst = {"norm": (0.0, ""), "proj": (0.0, ""), "full": (0.0, "")}
    n = 0
    for k, g in fx["grads"].items():
        if g["norm"] < 1e-6:
            continue
        got = params[k].grad.float().cpu()
        e_norm = abs(got.norm().item() - g["norm"]) / g["norm"]
        e_proj = abs((got * probe(got.shape, g["probe_seed"])).sum().item() - g["proj"]) / g["norm"]
        worst["norm"], worst["proj"] = max(worst["norm"], (e_norm, k)), max(worst["proj"], (e_proj, k))
        assert e_norm < 5e-2, (k, e_norm)
        assert e_proj < 12e-2, (k, e_proj)
        if g["full"] is not None:
            e_full = ((got - g["full"]).norm() / g["full"].norm()).item()
            worst["full"] = max(worst["full"], (e_full, k))
            assert e_full < 4e-2, (k, e_full)
        n += 1
    assert n > 90
    cb = m.visual_transformer.vq._codebook
    assert torch.allclose(cb.cluster_size.cpu(), fx["ema_cluster_size"], atol=1e-5)
    assert torch.allclose(cb.embed[0, :16].cpu(), fx["ema_embed_head"], atol=2e-3)
    # one optimiser step: replicas stay bit-identical (summed gradients are identical on every rank)
    tr.optimizer_step()
    chk = tr.arena.flat.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "replicas diverged after one optimiser step"
    tr.raise_if_skipped()
    mode = symm.mode()
    if mode == "peer" and symm.get_exchange(b, cfg["dim_latent"]) is None:
        mode = "nccl (peer mapping failed: fallback)"
    return f"loss {float(loss):.6f} (reference {float(fx['loss_train']):.6f}), {n} gradients, worst {worst}, exchange={mode}"


def kernel_mode(rank, world, b=8, d=512, steps=5):
    from ctpa_clip_b200 import ops, symm
    ex = symm.LatentExchange(b, d)
    dev = "cuda"
    for step in range(steps):
        g = torch.Generator().manual_seed(100 + step)
        T_all = torch.nn.functional.normalize(torch.randn(world * b, d, generator=g), dim=-1).to(dev)
        I_all = torch.nn.functional.normalize(torch.randn(world * b, d, generator=g), dim=-1).to(dev)
        tau = torch.tensor([0.3 + 0.2 * step], device=dev)
        sl = slice(rank * b, (rank + 1) * b)
        want = ops.clip_loss(T_all, I_all, tau, rank * b, b, want_grad=True)
        got = ops.clip_loss_allgather(T_all[sl].contiguous(), I_all[sl].contiguous(), tau, rank, world, ex.table,
                                      ex.next_step(), ex.status)
        torch.cuda.synchronize()
        for name, a, w in zip(("loss", "dT", "dI", "dtau"), got, want):
            if name == "dtau":
                assert torch.allclose(a, w, rtol=1e-5, atol=1e-7), (step, name)
            else:
                assert torch.equal(a, w), f"rank {rank} step {step}: {name} differs, max |d| = {(a - w).abs().max().item():.3e}"
    assert int(ex.status.item()) == 0
    dist.barrier()
    ex.close()
    return f"{steps} steps bit-exact"


def main():
    rank, world, port, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    assert torch.cuda.device_count() >= world, "one GPU per rank (never share a device between spinning ranks)"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ctpa_clip_b200 import symm
    symm.set_timeout(60.0)                                           # a test must fail, not hang, if a peer dies
    msg = train_mode(rank, world) if mode == "train" else kernel_mode(rank, world)
    dist.barrier()
    symm.shutdown()
    dist.destroy_process_group()
    print(f"DP_OK rank {rank}/{world} on cuda:{rank}: {msg}", flush=True)


if __name__ == "__main__":
    main()
