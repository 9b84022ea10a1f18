"""data_prep timing: batch of raw 512x512x320 int16 scans -> (240,480,480) fp32, CUDA events"""
import sys, json, torch
sys.path.insert(0, ".")
from ctpa_clip_b200.data_prep import preprocess_volumes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = torch.Generator(device="cuda").manual_seed(2)
raw = torch.randint(-1024, 3071, (B, 512, 512, 320), device="cuda", dtype=torch.int16, generator=g)
for icpt in (0.0, -1024.0):
    for _ in range(2):
        preprocess_volumes(raw, 1.0, icpt, 0.703125, 1.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        preprocess_volumes(raw, 1.0, icpt, 0.703125, 1.125)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byts = B * (512 * 512 * 320 * 2 + 240 * 480 * 480 * 4)
    print(json.dumps({"intercept": icpt, "batch": B, "ms": ms, "volumes_per_s": B / ms * 1e3, "algorithmic_GBps": byts / ms / 1e6,
                      "frac_of_6556": byts / ms / 1e6 / 6556.2}))
