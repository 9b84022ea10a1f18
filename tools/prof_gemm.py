"""single production-shape GEMM launches for ncu: ff1 forward (bf16 out), out-proj + residual (fp32 out)"""
import sys, torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops
T = 13824 * 8
g = torch.Generator(device="cuda").manual_seed(0)
xf = torch.randn(T, 512, device="cuda", generator=g).bfloat16()
w1 = torch.randn(2736, 512, device="cuda", generator=g).bfloat16()
for _ in range(2):
    h1 = ops.gemm(xf, w1)
q = torch.randn(T, 256, device="cuda", generator=g).bfloat16()
wo = torch.randn(512, 256, device="cuda", generator=g).bfloat16()
x = torch.randn(T, 512, device="cuda", generator=g)
for _ in range(2):
    x2 = ops.gemm(q, wo, out_dtype=torch.float32, resid=x)
torch.cuda.synchronize()
print("done")
