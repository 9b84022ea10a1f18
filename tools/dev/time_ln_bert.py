import sys, torch
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops
T, D = 4096, 768
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(T, D, device="cuda", generator=g); dy = torch.randn(T, D, device="cuda", generator=g)
gam = torch.randn(D, device="cuda", generator=g); dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
for _ in range(5): ops.layernorm_bwd(dy, x, gam, dgamma=dg, dbeta=db)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): ops.layernorm_bwd(dy, x, gam, dgamma=dg, dbeta=db)
b.record(); torch.cuda.synchronize()
print("BERT-shape layernorm_bwd [4096,768]: %.1f us per call" % (a.elapsed_time(b) * 1e3 / 50))
