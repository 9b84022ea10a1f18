"""Context number ("the bar is stock PyTorch running the same module on the same B200"):
the reference algorithm — the fp32 oracle port of the unmodified reference modules — run as it is on cuda:0 with stock
ATen / cuBLAS / cuDNN kernels, fp32 and under bf16 autocast, for the SAME training step bench.py times (production config,
forward + InfoNCE + backward + clip_grad_norm_(0.5) + Adam). Builder-run only; NOT a bench value and not on the product path.

    python tools/stock_torch_step.py [--batch 8] [--steps 5] [--warmup 2]      -> one JSON line per mode
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import ctclip_oracle as O  # noqa: E402  (test infrastructure, timed here as the "stock PyTorch" yardstick)


def run(batch, steps, warmup, autocast):
    cfg = O.PRODUCTION
    dev = torch.device("cuda", 0)
    sd = {k: v.to(dev) for k, v in O.init_state_dict(cfg, 0).items()}
    params = []
    for k, v in sd.items():
        if v.dtype.is_floating_point and v.numel() > 0 and "codebook" not in k and "beta" not in k:
            v.requires_grad_(True)
            params.append(v)
    txt = O.make_text_encoder(cfg, 0).to(dev)
    params += [p for n, p in txt.named_parameters() if not n.startswith("pooler.")]
    opt = torch.optim.Adam(params, lr=1.25e-6, betas=(0.9, 0.99), eps=1e-8, fused=True)
    video, ids, mask = (t.to(dev) for t in O.make_inputs(cfg, batch, 0))

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = O.ctclip_forward(sd, cfg, txt, ids, mask, video, training=True)["loss"]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"impl": "stock-torch (oracle port on cuda, ATen/cuBLAS/cuDNN)", "autocast_bf16": autocast, "batch": batch,
            "ms_per_step": ms, "volumes_per_s": batch / (ms * 1e-3), "loss": float(loss),
            "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    for autocast in (True, False):
        try:
            print(json.dumps(run(a.batch, a.steps, a.warmup, autocast)), flush=True)
        except torch.OutOfMemoryError as e:
            print(json.dumps({"impl": "stock-torch", "autocast_bf16": autocast, "batch": a.batch, "error": "OOM: " + str(e)[:120]}),
                  flush=True)
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
