"""one production-size launch of the data_prep kernel for ncu (2 scans)"""
import sys, torch
sys.path.insert(0, ".")
from ctpa_clip_b200.data_prep import preprocess_volumes
g = torch.Generator(device="cuda").manual_seed(2)
raw = torch.randint(-1024, 3071, (2, 512, 512, 320), device="cuda", dtype=torch.int16, generator=g)
for _ in range(2):
    preprocess_volumes(raw, 1.0, 0.0, 0.703125, 1.125)
torch.cuda.synchronize()
print("prep done")
