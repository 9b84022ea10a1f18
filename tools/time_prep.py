"""data_prep timing (BASELINE configs[3]): `n` raw int16 (512, 512, 320) scans -> (n, 240, 480, 480) fp32, CUDA events, resident
inputs (>> L2). Prints ms, algorithmic GB/s (389.0 MB per volume) and the fraction of the measured HBM copy bandwidth for the
kernel variants selected by CTCLIP_PREP_V2 / CTCLIP_PREP_V2_OCC / CTCLIP_PREP_X2 (one line each).
    python tools/time_prep.py [n=32]"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ctpa_clip_b200.data_prep import preprocess_volumes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
g = torch.Generator().manual_seed(2)
base = torch.randint(-1024, 3071, (4, 512, 512, 320), generator=g, dtype=torch.int16).cuda()
raw = base.repeat((n + 3) // 4, 1, 1, 1)[:n].contiguous()
out = torch.empty((n, 240, 480, 480), device="cuda")
BYTES = 512 * 512 * 320 * 2 + 240 * 480 * 480 * 4
variants = [(f"v2 nc{c} occ{o} tma{t}", {"CTCLIP_PREP_V2": "1", "CTCLIP_PREP_V2_NC": c, "CTCLIP_PREP_V2_OCC": o, "CTCLIP_PREP_V2_TMA": t})
            for c, o, t in (("1", "2", "1"), ("1", "2", "0"), ("1", "3", "1"), ("2", "2", "1"))] + [
    ("v1 x2", {"CTCLIP_PREP_V2": "0", "CTCLIP_PREP_X2": "1"})]
ref = None
for name, env in variants:
    os.environ.update(env)
    for _ in range(3):
        preprocess_volumes(raw, 1.0, 0.0, 0.703125, 1.125, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        preprocess_volumes(raw, 1.0, 0.0, 0.703125, 1.125, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    chk = out[:4].double().sum().item()
    ref = chk if ref is None else ref
    gbs = n * BYTES / (ms * 1e-3) / 1e9
    print(json.dumps({"variant": name, "volumes": n, "ms": round(ms, 4), "GBps_algorithmic": round(gbs, 1), "peak_GBps": peak,
                      "frac": round(gbs / peak, 4), "volumes_per_s": round(n / (ms * 1e-3), 1), "checksum_equal": chk == ref}))
