"""How much of a training step is the GPU waiting for the host to enqueue work? Two measurements of the same step:
   normal  : events around trainer.step() — what bench.py times
   pre-fed : a long device-side spin is enqueued first, so every launch of the step is already queued when the GPU starts
             on it; the step is timed from the end of the spin (event) to its last kernel
The difference is the launch-bound idle time a CUDA graph of the step could remove.
    python tools/host_starvation.py [batch=8]"""
import json, sys
import torch
sys.path.insert(0, ".")
from transformers import BatchEncoding
from ctpa_clip_b200.trainer import CTClipTrainStep
from ctpa_clip_b200 import configs as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = O.CONFIGS["production"]
dev = torch.device("cuda", 0)
model = O.build_model(cfg, dev)
trainer = CTClipTrainStep(model)
video, ids, mask = O.synth_batch(cfg, B, 100)
video = video.to(dev)
text = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})
for _ in range(3):
    trainer.step(text, video)
torch.cuda.synchronize()


def run(spin_cycles):
    ts = []
    for _ in range(6):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if spin_cycles:
            torch.cuda._sleep(spin_cycles)
        a.record()
        trainer.step(text, video)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


normal = run(0)
prefed = run(int(150e6))   # ~80 ms at 1.9 GHz: longer than the host needs to enqueue one step
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
torch.cuda._sleep(int(400e6)); trainer.step(text, video); t1 = time.perf_counter()
torch.cuda.synchronize()
print(json.dumps({"batch": B, "step_ms_normal": round(normal, 3), "step_ms_prefed": round(prefed, 3),
                  "host_bound_idle_ms": round(normal - prefed, 3), "host_enqueue_ms_per_step": round((t1 - t0) * 1e3, 2)}))
