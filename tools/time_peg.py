"""PEG kernels on the production token grid (B x 24 x 24 x 24 x 512, spatial + temporal call): CUDA-event time per call and
algorithmic GB/s; CTCLIP_PEG_STAGE=1 (cp.async staged planes, default) vs 0 (direct loads).  python tools/time_peg.py [B=8]"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ctpa_clip_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
t = h = w = 24
T, D = B * t * h * w, 512
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(T, D, device="cuda", generator=g)
dy = torch.randn(T, D, device="cuda", generator=g)
w27 = torch.randn(27, D, device="cuda", generator=g) / 5
bias = torch.randn(D, device="cuda", generator=g)
grid = (B, t, h, w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / iters


ref = {}
for stage in ("1", "0"):
    os.environ["CTCLIP_PEG_STAGE"] = stage
    for temporal in (False, True):
        dw, db = torch.zeros(27, D, device="cuda"), torch.zeros(D, device="cuda")
        outs = {"fwd": ops.peg_fwd(x, w27, bias, grid, temporal), "bwd_data": ops.peg_bwd_data(dy, w27, grid, temporal, want_bf16=True)[0]}
        ops.peg_bwd_weight(x, dy, dw, db, grid, temporal)
        outs["bwd_weight"] = dw
        cases = {"fwd": (lambda: ops.peg_fwd(x, w27, bias, grid, temporal), 2 * T * D * 4),
                 "bwd_data": (lambda: ops.peg_bwd_data(dy, w27, grid, temporal, want_bf16=True), 2 * T * D * 4 + T * D * 2),
                 "bwd_weight": (lambda: ops.peg_bwd_weight(x, dy, dw, db, grid, temporal), 2 * T * D * 4)}
        for name, (fn, nbytes) in cases.items():
            ms = timed(fn)
            key = (name, temporal)
            same = None
            if stage == "1":
                ref[key] = outs[name].clone()
            else:
                same = bool(torch.allclose(ref[key], outs[name], rtol=1e-4, atol=1e-4))
            print(json.dumps({"kernel": "peg_" + name, "temporal": temporal, "staged": stage == "1", "us": round(ms * 1e3, 1),
                              "GBps_algorithmic": round(nbytes / (ms * 1e-3) / 1e9, 1), "frac_of_hbm": round(nbytes / (ms * 1e-3) / 1e9 / peak, 3),
                              "equals_staged": same}))
