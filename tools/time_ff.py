"""FeedForward GEGLU fusions on the image tower's shape (T = 110592 tokens, dim 512, Nh = 1368): the GEMM with the GEGLU
epilogue beside the GEMM + row-wise kernel pair it replaces (cold L2 between calls).   python tools/time_ff.py"""
import json, sys
import torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops

T, D, NH = 110592, 512, 1368
g = torch.Generator(device="cuda").manual_seed(0)
xf = torch.randn(T, D, device="cuda", generator=g).bfloat16()
w1 = (torch.randn(2 * NH, D, device="cuda", generator=g) * 0.05).bfloat16()
w2 = (torch.randn(D, NH, device="cuda", generator=g) * 0.05).bfloat16()
gy = torch.randn(T, D, device="cuda", generator=g).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    for _ in range(3): fn()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ms += a.elapsed_time(b)
    return ms / iters * 1e3


h = ops.gemm(xf, w1)
out = {"ff1_gemm_us": timed(lambda: ops.gemm(xf, w1)), "geglu_fwd_us": timed(lambda: ops.geglu_fwd(h)),
       "ff1_gemm_geglu_fused_us": timed(lambda: ops.gemm_geglu(xf, w1))}
if hasattr(ops, "gemm_geglu_bwd"):
    du = ops.gemm(gy, w2, b_t=True)
    out.update({"ff2_dgrad_gemm_us": timed(lambda: ops.gemm(gy, w2, b_t=True)), "geglu_bwd_us": timed(lambda: ops.geglu_bwd(h, du)),
                "ff2_dgrad_geglu_fused_us": timed(lambda: ops.gemm_geglu_bwd(gy, w2, h))})
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
