import sys, torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops
T = 110592
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, iters=10):
    for _ in range(3): fn()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ms += a.elapsed_time(b)
    return ms / iters * 1e3
for K, N, name in ((512, 512, "dkv dgrad (g1 += dkv Wkv)"), (256, 512, "K=256"), (1368, 512, "ff2-like K=1368")):
    a = torch.randn(T, K, device="cuda", generator=g).bfloat16()
    w = torch.randn(N, K, device="cuda", generator=g).bfloat16()       # [N, K] row-major (K-major B, as wout / w2p)
    x = torch.randn(T, N, device="cuda", generator=g)
    y0 = torch.empty_like(x); ops.gemm(a, w, out=y0, resid=x)
    y1 = x.clone(); ops.gemm(a, w, out=y1, resid=y1)
    y2 = x.clone(); ops.gemm(a, w, out=y2, accumulate=True)
    print(name, "max diff", float((y1 - y2).abs().max()), float((y0 - y2).abs().max()), "resid out of place %.1f us" % timed(lambda: ops.gemm(a, w, out=y0, resid=x)), "| resid in place %.1f us" % timed(lambda: ops.gemm(a, w, out=y1, resid=y1)),
          "| red.add %.1f us" % timed(lambda: ops.gemm(a, w, out=y2, accumulate=True)),
          "| plain fp32 out %.1f us" % timed(lambda: ops.gemm(a, w, out=y2)),
          "| bf16 out %.1f us" % timed(lambda: ops.gemm(a, w)))
