"""LayerNorm backward / forward on the image tower's shape ([110592, 512] fp32 residual stream, bf16 upstream gradient,
fp32 residual-gradient add) against the measured HBM peak.   python tools/time_ln.py"""
import json, sys
import torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops

T, D = 110592, 512
PEAK = 6556.2
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(T, D, device="cuda", generator=g)
dy_bf = torch.randn(T, D, device="cuda", generator=g).bfloat16()
dy_f = torch.randn(T, D, device="cuda", generator=g)
add = torch.randn(T, D, device="cuda", generator=g)
gam = torch.randn(D, device="cuda", generator=g)
dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
out = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    for _ in range(3): fn()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ms += a.elapsed_time(b)
    return ms / iters * 1e3


n = T * D
for name, fn, nbytes in (
    ("ln_bwd bf16 dy + add_in -> fp32", lambda: ops.layernorm_bwd(dy_bf, x, gam, add_in=add, dgamma=dg, dbeta=db, out=out), n * (2 + 4 + 4 + 4)),
    ("ln_bwd bf16 dy -> fp32", lambda: ops.layernorm_bwd(dy_bf, x, gam, dgamma=dg, dbeta=db, out=out), n * (2 + 4 + 4)),
    ("ln_bwd fp32 dy + add_in -> fp32", lambda: ops.layernorm_bwd(dy_f, x, gam, add_in=add, dgamma=dg, dbeta=db, out=out), n * 16),
    ("ln_fwd fp32 -> bf16", lambda: ops.layernorm_fwd(x, gam, None, want_bf16=True), n * 6),
):
    us = timed(fn)
    print(json.dumps({"kernel": name, "us": round(us, 1), "GBps_algorithmic": round(nbytes / us / 1e3, 1), "frac_of_hbm": round(nbytes / us / 1e3 / PEAK, 3)}))
