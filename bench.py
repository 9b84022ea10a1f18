#!/usr/bin/env python
"""bench.py — CT volumes/sec of one CT-CLIP contrastive TRAINING step (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W            # this implementation (libctclip_sm100.so)
    torchrun ... bench.py --gpus N ...                       # one rank per GPU, NCCL over NVLink
    python bench.py --impl reference ...                     # the reference algorithm's CPU path (oracle port)

Workload (N=1): BASELINE.json configs[1] — production CT-CLIP (CTViT dim 512, patch 20x20x10, 4+4 layers, 8x32 heads,
codebook 8192, BERT-base text tower, 294912->512 latent projection), 8 synthetic 480x480x240 volumes + 8 reports of 512
token ids per rank; a step = forward(global-batch InfoNCE) + backward + gradient all-reduce + clip(0.5) + Adam.
N>1 keeps 8 volumes per rank (weak scaling; N=8 is configs[2], global batch 64; latents exchanged over NVLink by the loss
kernel itself, csrc/symm.cu). Prints ONE JSON line on rank 0: value = device-resident steps (CUDA events, max over ranks);
e2e = the same steps fed from pinned host memory (volumes, token ids, masks copied every step on a copy stream, every step's
loss copied back to pinned memory and consumed by the host one step later); gpu_launches = kernels of libctclip_sm100.so
launched by this rank inside the timed region (gpu_launches_per_step = per step); roofline / cpu_baseline as DESIGN.md §8.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "CT volumes/sec per CT-CLIP train step"
UNIT = "volumes/s"
TRAIN_GF_PER_VOLUME = 2428.0  # 887.4 GF forward + 1540.6 GF backward, un-padded dims


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="volumes per rank")
    ap.add_argument("--config", default="production")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default="", help="write the per-op device-time breakdown of one step to this file")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """samples SM clocks / throttle reasons with nvidia-smi while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1404.4), d.get("hbm_gbs", 6556.2), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_step(cfg, batch, threads, seed=0):
    """one forward+backward of the reference algorithm (oracle port, fp32) on the host cores; returns seconds"""
    import torch
    from oracle import ctclip_oracle as O
    torch.set_num_threads(threads)
    sd = O.init_state_dict(cfg, seed)
    sd = {k: v.requires_grad_(v.dtype.is_floating_point and v.numel() > 0 and "codebook" not in k and "beta" not in k)
          for k, v in sd.items()}
    txt = O.make_text_encoder(cfg, seed)
    video, ids, mask = O.make_inputs(cfg, batch, seed)
    t0 = time.perf_counter()
    out = O.ctclip_forward(sd, cfg, txt, ids, mask, video, training=True)
    out["loss"].backward()
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ctclip_oracle as O
    cfg = O.CONFIGS[args.config]
    threads = os.cpu_count() or 1
    sample_b = 1
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(cfg, sample_b, threads)
    times = [cpu_reference_step(cfg, sample_b, threads) for _ in range(max(1, min(args.steps, 8)))]
    t = sorted(times)[len(times) // 2]
    value = sample_b / t
    sample = f"{sample_b} volume(s) forward+backward per step, fp32, oracle port of the reference CT_CLIP on {threads} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CT-CLIP contrastive training step (forward + InfoNCE + backward), production config, CPU",
                   "volumes_per_step": sample_b},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from transformers import BatchEncoding
    from ctpa_clip_b200 import _lib, ops
    from ctpa_clip_b200.trainer import CTClipTrainStep
    from ctpa_clip_b200 import configs as O   # shapes + seeded synthetic inputs; the product arm imports nothing from oracle/

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = O.CONFIGS[args.config]
    B = args.batch
    model = O.build_model(cfg, dev, seed=0)
    trainer = CTClipTrainStep(model)
    video_h, ids, mask = O.synth_batch(cfg, B, seed=100 + rank)
    host = [video_h.pin_memory(), video_h.clone().pin_memory()]
    video_d = video_h.to(dev)
    text = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident steps
    def step_resident(i):
        trainer.step(text, video_d)

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ms_total = timed(step_resident, args.steps)
    launches_total = _lib.launch_count() - n0                      # kernels of libctclip_sm100.so inside the timed region (this rank)
    launches = launches_total // max(1, args.steps)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end: pinned host volumes -> (copy stream, double buffered) -> step -> loss read back
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(video_d), torch.empty_like(video_d)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    ids_h, mask_h = ids.pin_memory(), mask.pin_memory()
    tbufs = [(torch.empty_like(text.input_ids), torch.empty_like(text.attention_mask)) for _ in range(2)]
    texts = [BatchEncoding({"input_ids": a, "attention_mask": b}) for a, b in tbufs]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            bufs[i % 2].copy_(host[i % 2], non_blocking=True)          # this step's volumes ...
            tbufs[i % 2][0].copy_(ids_h, non_blocking=True)            # ... and token ids / attention mask
            tbufs[i % 2][1].copy_(mask_h, non_blocking=True)
            ready[i % 2].record(copy_stream)

    losses = []
    k = [0]                                   # running step index: step k reads bufs[k % 2], prefetches step k + 1
    loss_host = torch.empty(args.steps + 8, dtype=torch.float32).pin_memory()   # one pinned slot per step
    in_flight = []                            # (slot, event) of losses copied back but not yet consumed by the host

    def step_e2e(_):
        i = k[0]
        k[0] += 1
        prefetch(i + 1)                       # next step's volumes: H2D overlaps this step's kernels
        torch.cuda.current_stream().wait_event(ready[i % 2])
        loss = trainer.step(texts[i % 2], bufs[i % 2])
        consumed[i % 2].record()
        # D2H read of the step's result, EVERY step (the trainer's `loss.item()`, CTCLIPTrainer.py:346), as an async copy into
        # pinned memory; the host consumes step i-1's value here, while step i is already enqueued, so its Python prelude
        # (module walk, operand-cache rebuild) no longer idles the GPU after each step. The host never runs more than one
        # step ahead; the last value is consumed right after the closing synchronize.
        slot = i % loss_host.numel()
        loss_host[slot: slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        if in_flight:
            s_prev, e_prev = in_flight.pop(0)
            e_prev.synchronize()
            losses.append(float(loss_host[s_prev]))
        in_flight.append((slot, ev))

    for e in consumed:
        e.record()
    prefetch(0)
    for i in range(min(2, args.warmup)):      # the copy pipeline reaches steady state after two steps
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    for s_prev, e_prev in in_flight:          # timed() ended with a device synchronize: the last loss is already on the host
        e_prev.synchronize()
        losses.append(float(loss_host[s_prev]))
    in_flight.clear()
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * B / (ms_e2e * 1e-3)

    # ---- one instrumented step: per-op device time (CUDA events on the launching stream)
    ops.profile_begin()
    trainer.step(text, video_d)
    torch.cuda.synchronize()
    prof = ops.profile_end()
    gemm_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("gemm"))
    gemm_flops = sum(v["flops"] for k, v in prof.items() if k.startswith("gemm"))
    total_prof_ms = sum(v["ms"] for v in prof.values())
    tf_peak, hbm_peak, peak_src = peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    if args.breakdown and rank == 0:
        rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
        Path(args.breakdown).write_text(json.dumps({"ms_per_step": ms_step, "ops": rows}, indent=1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # DRAM bytes per GEMM launch: NOT measured in this run (a profiler is never attached to a bench run) — the committed ncu
    # metrics pass over one identical training step (tools/step_traffic.py -> profiles/r01_step_traffic.json)
    traffic, traffic_note = None, "profiles/r01_step_traffic.json not found"
    tpath = Path(__file__).resolve().parent / "profiles" / "r01_step_traffic.json"
    if tpath.exists():
        tj = json.loads(tpath.read_text())["gemm"]
        traffic = tj["dram_bytes_per_launch"]
        traffic_note = (f"bytes per launch: dram__bytes_read+write summed over the {tj['launches']} gemm_bf16_kernel launches of one "
                        f"step / launches, from the committed ncu pass profiles/r01_step_traffic.json "
                        f"(GEMM share of that serialised step: {tj['share_of_step']:.3f})")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "CT-CLIP contrastive training step (BASELINE.json configs[1]): production CTViT+BERT-base, "
                               f"{B} volumes 480x480x240 + {B} reports x 512 ids per rank, fwd+bwd+allreduce+clip+Adam",
                   "global_batch": world * B, "per_rank_batch": B, "parallelism": f"dp{world}",
                   "l2": "inputs larger than L2: 1.77 GB of volumes and >10 GB of activations per step vs 126 MB L2",
                   "text_tower": "BERT-base forward+backward on libctclip_sm100.so (tcgen05 GEMMs incl. batched per-head QK^T/PV, "
                                 "native softmax/GELU/LayerNorm/dropout kernels); HF module only holds the parameters"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(video_h.numel() * 4 + ids.numel() * 8 + mask.numel() * 8), "d2h_bytes_per_step": 4,
                "note": "pinned host volumes + token ids + masks, double-buffered copy stream; every step's loss is copied to pinned host memory and read by the host one step later (last one after the closing sync)"},
        "gpu_launches": int(launches_total), "gpu_launches_per_step": int(launches),
        "achieved_tflops_step": world * B * TRAIN_GF_PER_VOLUME / (ms_step * 1e-3) / 1e3 / world,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                     "frac": achieved / tf_peak if tf_peak else None, "traffic": traffic, "traffic_note": traffic_note,
                     "kernel": "gemm_bf16_kernel (tcgen05): all launches of one step, algorithmic (un-padded) FLOPs / "
                               "summed CUDA-event durations", "peak_source": peak_src,
                     "share_of_step": gemm_ms / total_prof_ms if total_prof_ms else None},
        "clocks": clocks,
        "loss": losses[-1] if losses else None,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        t = cpu_reference_step(cfg, 1, threads)
        out["cpu_baseline"] = {"value": 1.0 / t, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": "1 volume forward+backward, production config, fp32 oracle port, one run"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
