"""GPU dev check for the tcgen05 GEMM: all operand layouts vs an fp32 torch matmul of the same bf16 inputs."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops, _lib

torch.manual_seed(0)
dev = "cuda"
results = []

def run(*a, **k):
    try:
        return _run(*a, **k)
    except Exception as e:
        print("EXC", a, k, repr(e), flush=True)
        return False

def _run(M, N, K, a_t, b_t, out_dtype=torch.bfloat16, bias=False, resid=False, accumulate=False, splits=1, ldpad=0):
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    Ab, Bb = A.bfloat16(), B.bfloat16()
    ref = Ab.float() @ Bb.float().t()
    a_store = Ab.t().contiguous() if a_t else Ab
    b_store = Bb.t().contiguous() if b_t else Bb
    if ldpad:
        def pad(t):
            buf = torch.zeros(t.shape[0], (t.shape[1] + ldpad + 7) // 8 * 8, device=dev, dtype=t.dtype)
            buf[:, :t.shape[1]] = t
            return buf[:, :t.shape[1]]
        a_store, b_store = pad(a_store), pad(b_store)
    bias_t = torch.randn(N, device=dev) if bias else None
    resid_t = torch.randn(M, N, device=dev) if resid else None
    if bias: ref = ref + bias_t
    if resid: ref = ref + resid_t
    out = None
    if accumulate:
        out = torch.ones(M, N, device=dev)
        ref = ref + 1
    got = ops.gemm(a_store, b_store, a_t=a_t, b_t=b_t, out=out, out_dtype=out_dtype, bias=bias_t, resid=resid_t,
                   accumulate=accumulate, splits=splits)
    torch.cuda.synchronize()
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-2 if out_dtype == torch.bfloat16 else 2e-3) * max(scale, 1.0)
    ok = err <= tol
    r = dict(M=M, N=N, K=K, a_t=a_t, b_t=b_t, out=str(out_dtype), bias=bias, resid=resid, acc=accumulate,
             splits=splits, err=err, scale=scale, ok=ok)
    print(r, flush=True)
    results.append(r)
    return ok

ok = True
# simplest first: one tile, K-major both
ok &= run(128, 128, 64, False, False, torch.float32)
ok &= run(128, 256, 128, False, False, torch.float32)
for a_t in (False, True):
    for b_t in (False, True):
        ok &= run(256, 256, 256, a_t, b_t, torch.float32)
        ok &= run(296, 520, 200, a_t, b_t, torch.float32)          # ragged M,N,K
        ok &= run(1000, 96, 72, a_t, b_t, torch.bfloat16)           # BN=128 path, tiny K
ok &= run(13824, 512, 4000, False, False, torch.float32, bias=True)      # patch embed
ok &= run(13824, 2730, 512, False, False, torch.bfloat16)                # ff1
ok &= run(13824, 512, 1365 + 3, False, False, torch.float32, resid=True) # ff2 (K padded to 1368)
ok &= run(13824, 512, 2730, False, True, torch.bfloat16, ldpad=8)        # dgrad-like with padded ld
ok &= run(2736, 512, 13824, True, True, torch.float32, accumulate=True, splits=0)  # wgrad split-K
ok &= run(512, 256, 13824, True, True, torch.float32, accumulate=True, splits=4)

# timing of the big shapes
def bench(M, N, K, a_t=False, b_t=False, out_dtype=torch.bfloat16, accumulate=False, splits=1, iters=20):
    A = torch.randn(K, M, device=dev).bfloat16() if a_t else torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(K, N, device=dev).bfloat16() if b_t else torch.randn(N, K, device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if accumulate else out_dtype)
    for _ in range(3):
        ops.gemm(A, B, a_t=a_t, b_t=b_t, out=out, accumulate=accumulate, splits=splits)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(A, B, a_t=a_t, b_t=b_t, out=out, accumulate=accumulate, splits=splits)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # torch reference timing
    if not a_t and not b_t:
        for _ in range(3): torch.matmul(A, B.t())
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters): torch.matmul(A, B.t())
        e1.record(); torch.cuda.synchronize()
        tms = e0.elapsed_time(e1) / iters
    else:
        tms = float("nan")
    r = dict(bench=True, M=M, N=N, K=K, a_t=a_t, b_t=b_t, ms=ms, tflops=tf, torch_ms=tms,
             torch_tflops=2.0 * M * N * K / tms / 1e9 if tms == tms else None)
    print(r, flush=True)
    results.append(r)

T = 13824 * 8
bench(T, 512, 4000)
bench(T, 2730, 512)
bench(T, 512, 1368, out_dtype=torch.float32)
bench(T, 512, 512)
bench(T, 8192, 512)
bench(T, 512, 2730, b_t=True)
bench(2736, 512, T, a_t=True, b_t=True, accumulate=True, splits=0)
bench(8192, 8192, 8192)
print("ALL_OK" if ok else "SOME_FAILED")
json.dump(results, open("gpurun_out/gemm_check.json", "w"), indent=1)
sys.exit(0 if ok else 1)
