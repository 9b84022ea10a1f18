"""Spatial attention forward / backward at the production shape (8 volumes: 192 frames x 576 tokens x 8 heads x 32),
CUDA events, cold L2.   python tools/time_attn.py [batch=8]"""
import json, os, sys
import torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops
from ctpa_clip_b200.ct_clip.attention import pair_index

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
grid, heads = (B, 24, 24, 24), 8
T, inner = B * 24 * 576, heads * 32
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(T, inner, device="cuda", generator=g).bfloat16()
kv = torch.randn(T, 2 * inner, device="cuda", generator=g).bfloat16()
qs = 1 + 0.1 * torch.randn(32, device="cuda", generator=g)
ks = 1 + 0.1 * torch.randn(32, device="cuda", generator=g)
tab = torch.randn(heads, 47 * 47, device="cuda", generator=g)
rowmax = tab[:, pair_index(24, 24, "cuda")].amax(dim=-1).contiguous()
d_o = torch.randn(T, inner, device="cuda", generator=g).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
o, lse = ops.attn_fwd(q, kv, grid, heads, False, qs, ks, tab, rowmax)
dqs, dks, dtab = torch.zeros(32, device="cuda"), torch.zeros(32, device="cuda"), torch.zeros_like(tab)


def timed(fn, iters=10):
    for _ in range(3): fn()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ms += a.elapsed_time(b)
    return ms / iters * 1e3


flops_fwd = 4.0 * B * 24 * heads * 576 * 576 * 32
out = {"batch": B, "attn_fwd_us": timed(lambda: ops.attn_fwd(q, kv, grid, heads, False, qs, ks, tab, rowmax))}
out["attn_bwd_us"] = timed(lambda: ops.attn_bwd(q, kv, o, lse, d_o, grid, heads, False, qs, ks, dqs, dks, tab, rowmax, dtab))
out = {k: round(v, 1) if isinstance(v, float) else v for k, v in out.items()}
out["fwd_TFLOPs"] = round(flops_fwd / out["attn_fwd_us"] / 1e6, 1)
out["bwd_TFLOPs"] = round(2.5 * flops_fwd / out["attn_bwd_us"] / 1e6, 1)
print(json.dumps(out))
