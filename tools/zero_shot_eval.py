"""BASELINE.json configs[4]: zero-shot CT-CLIP inference, 256 synthetic volumes x 18 pathology prompt pairs, replicas over N GPUs.
    python tools/zero_shot_eval.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/zero_shot_eval.py
Every rank scores its round-robin shard of the volumes with `ctpa_clip_b200.inference.ZeroShotEvaluator` (prompt latents
once, each volume through the image tower once, `ctclip_zero_shot_scores`), the (256, 18) matrix is gathered; rank 0 prints
one JSON line (volumes/s over all ranks, device time, max over ranks). Volumes are generated on the host and copied per batch
(the reference's DataLoader hands CPU tensors, ctclip_inference.py:305-310)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist
from transformers import BatchEncoding


class SyntheticVolumes:
    """`n` seeded volumes (1, f, h, w) in [-1, 1]; a few distinct ones are cycled so that host memory stays small"""

    def __init__(self, cfg, n, distinct=8):
        g = torch.Generator().manual_seed(5)
        self.base = [(torch.rand(1, cfg["frames"], cfg["image_size"], cfg["image_size"], generator=g) * 2 - 1).pin_memory()
                     for _ in range(distinct)]
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.base[i % len(self.base)]


def main():
    import bench
    from ctpa_clip_b200.inference import PATHOLOGIES, ZeroShotEvaluator
    from ctpa_clip_b200 import configs as O
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = O.CONFIGS[os.environ.get("ZS_CONFIG", "production")]
    n_vol = int(os.environ.get("ZS_VOLUMES", "256"))
    model = O.build_model(cfg, dev, seed=0).eval()
    # no tokenizer offline: 36 seeded id rows stand for the tokenised "<pathology> is present." / "... is not present." pairs
    g = torch.Generator().manual_seed(9)
    P, L = len(PATHOLOGIES), cfg["seq_len"]
    ids = torch.randint(1, cfg["text"]["vocab_size"], (2 * P, L), generator=g)
    mask = torch.ones(2 * P, L, dtype=torch.long)
    ids[:, 16:] = 0
    mask[:, 16:] = 0
    prompts = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})
    data = SyntheticVolumes(cfg, n_vol)
    ev = ZeroShotEvaluator(model, prompts, batch_size=8)
    ev.evaluate(SyntheticVolumes(cfg, 8 * world))                 # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pred, _ = ev.evaluate(data)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "zero-shot volumes/s (256 volumes x 18 prompt pairs)", "value": n_vol / (float(ms) * 1e-3),
                          "unit": "volumes/s", "n_gpus": world, "ms_total": float(ms), "scores_shape": list(pred.shape),
                          "finite": bool((pred == pred).all()), "data": "synthetic", "parallelism": f"replicas x{world}"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
