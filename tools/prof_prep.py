"""ncu target: a few launches of the data_prep kernel on 8 production-size scans (python tools/prof_prep.py [n=8])"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ctpa_clip_b200.data_prep import preprocess_volumes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = torch.Generator().manual_seed(2)
raw = torch.randint(-1024, 3071, (n, 512, 512, 320), generator=g, dtype=torch.int16).cuda()
out = torch.empty((n, 240, 480, 480), device="cuda")
for _ in range(3):
    preprocess_volumes(raw, 1.0, 0.0, 0.703125, 1.125, out=out)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0, 0]))
