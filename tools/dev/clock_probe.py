"""SM clock / board power while one GEMM variant runs back to back (is the epilogue cost a power-cap effect?)"""
import os, sys, threading, time, json
import torch, pynvml
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops
pynvml.nvmlInit(); hd = pynvml.nvmlDeviceGetHandleByIndex(0)
T, D, NH = 110592, 512, 1368
g = torch.Generator(device="cuda").manual_seed(0)
xf = torch.randn(T, D, device="cuda", generator=g).bfloat16()
w1 = (torch.randn(2 * NH, D, device="cuda", generator=g) * 0.05).bfloat16()
h = ops.gemm(xf, w1)
def probe(name, fn, secs=1.5):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    samples, stop = [], False
    def sampler():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1e3))
            time.sleep(0.02)
    th = threading.Thread(target=sampler); th.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); a.record()
    while time.time() - t0 < secs:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize(); stop = True; th.join()
    s = samples[len(samples) // 3:]
    print(json.dumps({"variant": name, "us": round(a.elapsed_time(b) * 1e3 / n, 1), "sm_mhz_median": sorted(x[0] for x in s)[len(s) // 2],
                      "power_w_median": round(sorted(x[1] for x in s)[len(s) // 2])}))
probe("plain gemm", lambda: ops.gemm(xf, w1))
for d in (7, 6, 1, 0):
    os.environ["CTCLIP_GEGLU_DBG"] = str(d)
    probe(f"fused dbg={d}", lambda: ops.gemm_geglu(xf, w1))
probe("geglu_fwd rowwise", lambda: ops.geglu_fwd(h))
