"""Component measurements that complement bench.py (SURVEY §8(d) configs 4 and 5 + kernel rooflines).
Writes one JSON document; run on a B200:  python tools/bench_components.py gpurun_out/components.json"""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops
from ctpa_clip_b200.data_prep import preprocess_volumes

dev = "cuda"
out = {}
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

# ---- config 4: data_prep, batch 32 raw 512x512x320 int16 -> (240,480,480) fp32
B = 32
g = torch.Generator(device=dev).manual_seed(2)
raw = torch.randint(-1024, 3071, (B, 512, 512, 320), device=dev, dtype=torch.int16, generator=g)
ms = timeit(lambda: preprocess_volumes(raw, 1.0, 0.0, 0.703125, 1.125), iters=5, warm=2)
algo_bytes = B * (512 * 512 * 320 * 2 + 240 * 480 * 480 * 4)
out["data_prep_batch32"] = {"ms": ms, "volumes_per_s": B / (ms * 1e-3), "algorithmic_GBps": algo_bytes / (ms * 1e-3) / 1e9,
                            "hbm_peak_GBps": peaks["hbm_gbs"], "frac_of_measured_hbm": algo_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                            "bytes_per_volume": algo_bytes // B}
# host -> device -> prep -> (stays on device): e2e with pinned host scans
host = raw[:8].cpu().pin_memory()
def e2e():
    d = host.to(dev, non_blocking=True)
    preprocess_volumes(d, 1.0, 0.0, 0.703125, 1.125)
ms8 = timeit(e2e, iters=3, warm=1)
out["data_prep_e2e_8_pinned_host"] = {"ms": ms8, "volumes_per_s": 8 / (ms8 * 1e-3), "h2d_bytes": int(host.numel() * 2)}
del raw
torch.cuda.empty_cache()
# CPU baseline: C oracle, one volume, one thread
from oracle import resample_oracle as R
rng = np.random.default_rng(2)
raw1 = rng.integers(-1024, 3071, size=(512, 512, 320), dtype=np.int16)
t0 = time.perf_counter(); R.preprocess_volume(raw1, 1.0, 0.0, 0.703125, 1.125); t = time.perf_counter() - t0
out["data_prep_cpu_oracle_1thread"] = {"s_per_volume": t, "volumes_per_s": 1 / t, "kind": "port (oracle/resample_oracle.c)"}

# ---- kernel rooflines on production shapes (B=8)
T = 13824 * 8
def gemm_case(name, M, N, K, **kw):
    A = torch.randn(K, M, device=dev).bfloat16() if kw.get("a_t") else torch.randn(M, K, device=dev).bfloat16()
    Bm = torch.randn(K, N, device=dev).bfloat16() if kw.get("b_t") else torch.randn(N, K, device=dev).bfloat16()
    acc = kw.pop("accumulate", False)
    o = torch.zeros(M, N, device=dev, dtype=torch.float32 if acc or kw.get("f32") else torch.bfloat16)
    kw.pop("f32", None)
    ms = timeit(lambda: ops.gemm(A, Bm, out=o, accumulate=acc, splits=0 if acc else 1, **kw), iters=20)
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    out["gemm_" + name] = {"ms": ms, "TFLOPs": tf, "frac_of_measured_bf16_burst": tf / peaks["bf16_tflops"]}
gemm_case("patch_embed_T x512x4000", T, 512, 4000, f32=True)
gemm_case("ff1_Tx2736x512", T, 2736, 512)
gemm_case("ff1_dgrad_Tx512x2736", T, 512, 2736, b_t=True, f32=True)
gemm_case("ff1_wgrad_2736x512xT", 2736, 512, T, a_t=True, b_t=True, accumulate=True)
gemm_case("vq_scores_Tx8192x512", T, 8192, 512)
gemm_case("square_8192", 8192, 8192, 8192)
# memory-bound projections: fp32 output added in place to the fp32 residual stream (out-proj, attention.py:181 + :326)
qo = torch.randn(T, 256, device=dev).bfloat16(); wo = torch.randn(512, 256, device=dev).bfloat16(); xr = torch.randn(T, 512, device=dev)
ms = timeit(lambda: ops.gemm(qo, wo, out=xr, resid=xr), iters=20)
byts = T * 256 * 2 + 2 * T * 512 * 4
out["gemm_outproj_resid_Tx512x256"] = {"ms": ms, "algorithmic_GBps": byts / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": byts / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
xk = torch.randn(T, 512, device=dev).bfloat16(); wk = torch.randn(512, 512, device=dev).bfloat16()
ms = timeit(lambda: ops.gemm(xk, wk), iters=20)
byts = 2 * T * 512 * 2
out["gemm_kv_Tx512x512_bf16"] = {"ms": ms, "algorithmic_GBps": byts / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": byts / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
x = torch.randn(T, 512, device=dev)
w27 = torch.randn(27, 512, device=dev); bias = torch.randn(512, device=dev)
ms = timeit(lambda: ops.peg_fwd(x, w27, bias, (8, 24, 24, 24), False))
out["peg_fwd"] = {"ms": ms, "algorithmic_GBps": 2 * x.numel() * 4 / (ms * 1e-3) / 1e9}
gam = torch.ones(512, device=dev)
ms = timeit(lambda: ops.layernorm_fwd(x, gam, None, want_bf16=True, want_raw_bf16=True))
out["layernorm_fwd_dual"] = {"ms": ms, "algorithmic_GBps": (x.numel() * 4 + 2 * x.numel() * 2) / (ms * 1e-3) / 1e9}
vid = torch.rand(8, 1, 240, 480, 480, device=dev)
g4, b4 = torch.ones(4000, device=dev), torch.zeros(4000, device=dev)
ms = timeit(lambda: ops.patch_ln_fwd(vid, g4, b4, 10, 20), iters=5)
out["patch_ln_fwd"] = {"ms": ms, "algorithmic_GBps": (vid.numel() * 4 + T * 4000 * 2) / (ms * 1e-3) / 1e9}
# ---- config 5: zero-shot scoring, one GPU's share (32 of the 256 volumes) x 18 pathologies x 2 prompts
del vid
torch.cuda.empty_cache()
sys.argv = sys.argv[:1]
import bench
from transformers import BatchEncoding
from ctpa_clip_b200 import configs as O
cfg = O.CONFIGS["production"]
model = O.build_model(cfg, torch.device("cuda"), seed=0).eval()
vols = torch.rand(32, 1, 240, 480, 480, device=dev) * 2 - 1
gq = torch.Generator().manual_seed(5)
pid = torch.randint(1, 30522, (36, 512), generator=gq); pmask = torch.ones(36, 512, dtype=torch.long)
pid[:, 32:] = 0; pmask[:, 32:] = 0
prompts = BatchEncoding({"input_ids": pid.to(dev), "attention_mask": pmask.to(dev)})
def zs():
    outs = [model.zero_shot_scores(prompts, vols[i:i + 8]) for i in range(0, 32, 8)]
    return torch.cat(outs)
ms = timeit(zs, iters=3, warm=1)
out["zero_shot_32vol_x_18path"] = {"ms": ms, "volumes_per_s": 32 / (ms * 1e-3), "scores_shape": list(zs().shape),
                                   "note": "image encoder once per volume (reference: 18x), text latents recomputed per 8-volume chunk"}
for k, v in out.items():
    print(k, json.dumps(v))
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/components.json", "w"), indent=1)
