"""fused BERT attention timing: B=8, H=12, L=512, D=768, half of the keys padding (the bench's report batch)"""
import sys, torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops
B, H, L, D = 8, 12, 512, 768
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 3 * D, device="cuda", generator=g).bfloat16()
dout = torch.randn(B * L, D, device="cuda", generator=g).bfloat16()
def t(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for pad in (256, 0):
    mask = torch.ones(B, L, dtype=torch.long, device="cuda")
    if pad: mask[:, L - pad:] = 0
    for p in (0.0, 0.1):
        cv, lse = ops.bert_attn_fwd(qkv, mask, B, H, L, D, p, 1)
        f = t(lambda: ops.bert_attn_fwd(qkv, mask, B, H, L, D, p, 1))
        b = t(lambda: ops.bert_attn_bwd(qkv, mask, cv, lse, dout, B, H, L, D, p, 1))
        print(f"pad {pad} dropout {p}: fwd {f:.1f} us  bwd {b:.1f} us")
