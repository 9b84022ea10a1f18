"""Kineto (CUPTI) timeline of one production training step: per-kernel device time, GPU busy vs idle inside the step.
Not a benchmark (profiler attached) — it answers "where does the step go besides our kernels".
    python tools/timeline.py gpurun_out/timeline.json [batch]"""
import json, sys
import torch
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
from transformers import BatchEncoding
import bench
from ctpa_clip_b200.trainer import CTClipTrainStep
from ctpa_clip_b200 import configs as O

B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = O.CONFIGS["production"]
dev = torch.device("cuda", 0)
model = O.build_model(cfg, dev)
trainer = CTClipTrainStep(model)
video, ids, mask = O.synth_batch(cfg, B, 100)
video = video.to(dev)
text = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})
for _ in range(2):
    trainer.step(text, video)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    trainer.step(text, video)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda x: x[0])
t0, t1 = ks[0][0], max(k[1] for k in ks)
busy, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
gaps = []
for s, e, n in ks[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, n))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
agg = {}
for s, e, n in ks:
    a = agg.setdefault(n[:90], [0.0, 0])
    a[0] += e - s
    a[1] += 1
ours = sum(v[0] for k, v in agg.items() if "<unnamed>" in k or "ctclip" in k)
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
out = {"span_ms": (t1 - t0) / 1e3, "busy_ms": busy / 1e3, "idle_ms": (t1 - t0 - busy) / 1e3, "our_kernels_ms": ours / 1e3,
       "other_kernels_ms": (sum(v[0] for v in agg.values()) - ours) / 1e3, "n_kernels": len(ks),
       "largest_gaps_us": [(round(g, 1), n[:60]) for g, n in sorted(gaps, key=lambda x: -x[0])[:15]],
       "top": [(k, round(v[0] / 1e3, 3), v[1]) for k, v in rows[:60]]}
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json", "w"), indent=1)
print(json.dumps({k: out[k] for k in ("span_ms", "busy_ms", "idle_ms", "our_kernels_ms", "other_kernels_ms", "n_kernels")}))
