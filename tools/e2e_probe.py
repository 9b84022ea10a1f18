"""Where does the end-to-end gap come from? Times the production step (B=8) in four ways:
resident / resident + loss.item() per step / pinned-host copies without the per-step read / both (= bench.py e2e)."""
import sys, json, torch
sys.path.insert(0, ".")
sys.argv = sys.argv[:1]
import bench
from transformers import BatchEncoding
from ctpa_clip_b200.trainer import CTClipTrainStep
from ctpa_clip_b200 import configs as O
cfg = O.CONFIGS["production"]
dev = torch.device("cuda", 0)
model = O.build_model(cfg, dev)
tr = CTClipTrainStep(model)
video_h, ids, mask = O.synth_batch(cfg, 8, 100)
host = [video_h.pin_memory(), video_h.clone().pin_memory()]
video_d = video_h.to(dev)
text = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})
copy_stream = torch.cuda.Stream()
bufs = [torch.empty_like(video_d), torch.empty_like(video_d)]
ready = [torch.cuda.Event(), torch.cuda.Event()]
consumed = [torch.cuda.Event(), torch.cuda.Event()]

def run(copies, sync, steps=6):
    for e in consumed: e.record()
    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            bufs[i % 2].copy_(host[i % 2], non_blocking=True)
            ready[i % 2].record(copy_stream)
    k = [0]
    if copies: prefetch(0)
    def step(_):
        i = k[0]; k[0] += 1
        if copies:
            prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            loss = tr.step(text, bufs[i % 2])
            consumed[i % 2].record()
        else:
            loss = tr.step(text, video_d)
        if sync: float(loss.detach())
    step(0); step(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): step(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
out = {}
for name, c, s in (("resident", 0, 0), ("resident+item", 0, 1), ("copies", 1, 0), ("copies+item", 1, 1)):
    out[name] = run(c, s)
print(json.dumps(out))
