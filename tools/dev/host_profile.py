"""cProfile of the host side of a training step (where do the ~19 ms of enqueue time go?)"""
import cProfile, pstats, sys, io
import torch
sys.path.insert(0, ".")
from transformers import BatchEncoding
from ctpa_clip_b200.trainer import CTClipTrainStep
from ctpa_clip_b200 import configs as O
cfg = O.CONFIGS["production"]
dev = torch.device("cuda", 0)
model = O.build_model(cfg, dev)
trainer = CTClipTrainStep(model)
video, ids, mask = O.synth_batch(cfg, 8, 100)
video = video.to(dev)
text = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})
for _ in range(3):
    trainer.step(text, video)
torch.cuda.synchronize()
torch.cuda._sleep(int(600e6))     # the device is busy: the host never waits for it
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    trainer.step(text, video)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30)
print(s.getvalue()[:6000])
