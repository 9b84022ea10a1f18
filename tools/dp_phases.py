"""Where a data-parallel training step spends its time, per phase and per rank (CUDA events on each rank's stream):
forward (both towers + latent exchange + loss) | backward (incl. the all-reduces / gathers issued from inside it) |
gradient reduction still exposed after backward | clip + Adam. Rank 0 prints one JSON line: per phase the min / median / max
over ranks of the mean over the timed steps, plus the barrier-bracketed step time. Run it at N = 1 and N = 8 and subtract.

    python tools/dp_phases.py                                                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/dp_phases.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist
from transformers import BatchEncoding


def main():
    from ctpa_clip_b200 import configs as O
    from ctpa_clip_b200.trainer import CTClipTrainStep
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = O.CONFIGS["production"]
    B, steps = int(os.environ.get("PHASE_BATCH", "8")), int(os.environ.get("PHASE_STEPS", "8"))
    model = O.build_model(cfg, dev, seed=0)
    tr = CTClipTrainStep(model)
    video, ids, mask = O.synth_batch(cfg, B, seed=100 + rank)
    video = video.to(dev)
    text = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})
    for _ in range(3):
        tr.step(text, video)
    names = ["forward", "backward", "reduce_exposed", "optimizer"]
    acc = torch.zeros(len(names) + 1, dtype=torch.float64)
    for _ in range(steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        if not model.training:
            model.train()
        loss = model(text, video, return_loss=True)
        ev[1].record()
        loss.backward()
        ev[2].record()
        tr.reduce_gradients()
        ev[3].record()
        tr.optimizer_step()
        ev[4].record()
        torch.cuda.synchronize()
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
        acc[4] += ev[0].elapsed_time(ev[4])
    acc /= steps
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allr, acc.tolist())
    else:
        allr = [acc.tolist()]
    if rank == 0:
        t = torch.tensor(allr)
        out = {"n_gpus": world, "per_rank_batch": B, "steps": steps, "unit": "ms"}
        for i, n in enumerate(names + ["step"]):
            col = t[:, i]
            out[n] = {"min": round(float(col.min()), 3), "median": round(float(col.median()), 3), "max": round(float(col.max()), 3)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
