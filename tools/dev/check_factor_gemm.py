"""The weight-gradient GEMM of the factor gather at the global batches of 2 / 4 / 8 ranks (K = 16, 32, 64):
dW[512, 294912] += dy_all^T x_all with MN-major bf16 operands and the in-place fp32 residual epilogue, against torch fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ctpa_clip_b200 import ops

torch.backends.cuda.matmul.allow_tf32 = False
ok = True
for K in (8, 16, 32, 64):
    g = torch.Generator(device="cuda").manual_seed(K)
    dy = torch.randn(K, 512, device="cuda", generator=g).bfloat16()
    x = torch.randn(K, 294912, device="cuda", generator=g).bfloat16()
    w = torch.randn(512, 294912, device="cuda", generator=g)
    ref = w + dy.float().t() @ x.float()
    ops.gemm(dy, x, a_t=True, b_t=True, out=w, resid=w)
    torch.cuda.synchronize()
    err = (w - ref).abs().max().item()
    print(f"K={K}: max |err| = {err:.3e} (scale {ref.abs().max().item():.2f})", flush=True)
    ok &= err < 1e-3 * ref.abs().max().item()
print("FACTOR_GEMM_OK" if ok else "FACTOR_GEMM_FAIL")
sys.exit(0 if ok else 1)
