"""our tcgen05 GEMM beside torch.matmul (cuBLAS) on the step's shapes — a yardstick only, cuBLAS is not on the product path.
   python tools/gemm_vs_cublas.py > gpurun_out/gemm_vs_cublas.txt"""
import sys, torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3, cold=False):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    tot = 0.0
    if not cold:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


T = 13824 * 8
cases = [  # name, M, N, K, a_t, b_t, f32 out
    ("bert_qkv", 4096, 2304, 768, False, False, False),
    ("bert_out", 4096, 768, 768, False, False, False),
    ("bert_ff1", 4096, 3072, 768, False, False, False),
    ("bert_ff2", 4096, 768, 3072, False, False, False),
    ("bert_dgrad_ff1", 4096, 768, 3072, False, True, False),
    ("bert_dgrad_qkv", 4096, 768, 2304, False, True, False),
    ("bert_wgrad_qkv", 2304, 768, 4096, True, True, True),
    ("bert_wgrad_ff", 3072, 768, 4096, True, True, True),
    ("bert_wgrad_out", 768, 768, 4096, True, True, True),
    ("vit_q", T, 256, 512, False, False, False),
    ("vit_kv", T, 512, 512, False, False, False),
    ("vit_out", T, 512, 256, False, False, True),
    ("vit_ff1", T, 2736, 512, False, False, False),
    ("vit_ff2", T, 512, 1368, False, False, True),
    ("vit_ff2_dgrad", T, 1368, 512, False, True, False),
    ("vit_ff1_dgrad", T, 512, 2736, False, True, True),
    ("vit_ff1_wgrad", 2736, 512, T, True, True, True),
    ("vit_ff2_wgrad", 512, 1368, T, True, True, True),
    ("vit_kv_wgrad", 512, 512, T, True, True, True),
    ("vit_q_wgrad", 256, 512, T, True, True, True),
    ("vit_out_wgrad", 512, 256, T, True, True, True),
    ("latent_fwd", 8, 512, 294912, False, False, True),
]
print(f"{'case':18s} {'M':>7s} {'N':>5s} {'K':>7s}   ours_us  cublas_us  ours_TF  cublas_TF   cold: ours cublas")
for name, M, N, K, a_t, b_t, f32 in cases:
    A = (torch.randn(K, M, device=dev) if a_t else torch.randn(M, K, device=dev)).bfloat16()
    B = (torch.randn(K, N, device=dev) if b_t else torch.randn(N, K, device=dev)).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    ours = lambda: ops.gemm(A, B, a_t=a_t, b_t=b_t, out=out, accumulate=a_t or M == 8, splits=0 if (a_t or M == 8) else 1)
    Am = A.t() if a_t else A
    Bm = B if b_t else B.t()
    ref_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ref = lambda: torch.matmul(Am, Bm, out=ref_out)
    t0, t1 = timeit(ours), timeit(ref)
    c0, c1 = timeit(ours, iters=5, cold=True), timeit(ref, iters=5, cold=True)
    fl = 2.0 * M * N * K
    print(f"{name:18s} {M:7d} {N:5d} {K:7d}  {t0:8.1f}  {t1:9.1f}  {fl/t0/1e6:7.0f}  {fl/t1/1e6:9.0f}   {c0:8.1f} {c1:8.1f}")
