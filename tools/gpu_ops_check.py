"""GPU dev check of the memory-bound kernels and the attention forward against torch fp32 references."""
import sys, json
import torch, torch.nn.functional as F
sys.path.insert(0, ".")
from ctpa_clip_b200 import ops
from oracle import ctclip_oracle as O
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
allok = True
def report(name, got, ref, tol):
    global allok
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= tol * max(scale, 1e-6)
    allok &= ok
    print(f"{name:40s} err={err:.3e} scale={scale:.3e} {'OK' if ok else 'FAIL'}", flush=True)

def guard(fn):
    global allok
    try:
        fn()
    except Exception as e:
        allok = False
        print("EXC", fn.__name__, repr(e), flush=True)

def t_layernorm():
    for rows, dim in [(1000, 512), (77, 64), (300, 128)]:
        x = torch.randn(rows, dim, device=dev) * 2 + 0.5
        g = 1 + 0.1 * torch.randn(dim, device=dev); b = 0.1 * torch.randn(dim, device=dev)
        y, raw, yf = ops.layernorm_fwd(x, g, b, want_bf16=True, want_raw_bf16=True, want_f32=True)
        ref = F.layer_norm(x, (dim,), g, b)
        report(f"ln_fwd f32 {rows}x{dim}", yf, ref, 1e-5); report("ln_fwd bf16", y, ref, 1e-2); report("ln_fwd raw", raw, x, 1e-2)
        y2, _, _ = ops.layernorm_fwd(x, g, None)
        report("ln_fwd gamma-only", y2, F.layer_norm(x, (dim,), g, None), 1e-2)
        xr = x.clone().requires_grad_(); gr = g.clone().requires_grad_(); br = b.clone().requires_grad_()
        dy = torch.randn(rows, dim, device=dev)
        F.layer_norm(xr, (dim,), gr, br).backward(dy)
        add = torch.randn(rows, dim, device=dev)
        dg = torch.zeros(dim, device=dev); db = torch.zeros(dim, device=dev)
        dx, dxb = ops.layernorm_bwd(dy, x, g, add_in=add, dgamma=dg, dbeta=db, want_bf16=True)
        report("ln_bwd dx", dx, xr.grad + add, 1e-4); report("ln_bwd dx_bf16", dxb, xr.grad + add, 1e-2)
        report("ln_bwd dgamma", dg, gr.grad, 1e-4); report("ln_bwd dbeta", db, br.grad, 1e-4)

def t_geglu():
    rows, half = 513, 1368
    h = torch.randn(rows, 2 * half, device=dev).bfloat16()
    u = ops.geglu_fwd(h)
    hf = h.float().requires_grad_()
    ref = hf[:, :half] * F.gelu(hf[:, half:])
    report("geglu_fwd", u, ref, 1e-2)
    du = torch.randn(rows, half, device=dev).bfloat16()
    ref.backward(du.float())
    report("geglu_bwd", ops.geglu_bwd(h, du), hf.grad, 1e-2)
    x = torch.randn(1000, 512, device=dev)
    report("cast", ops.cast_bf16(x), x, 1e-2)

def t_peg():
    for (b, t, h, w, dim) in [(2, 5, 4, 4, 64), (2, 6, 6, 6, 128), (1, 24, 24, 24, 512)]:
        sd = {"dsconv.weight": (torch.randn(dim, 1, 3, 3, 3, device=dev) / 5), "dsconv.bias": torch.randn(dim, device=dev) * 0.1}
        w27 = sd["dsconv.weight"].reshape(dim, 27).t().contiguous()
        for temporal in (False, True):
            xc = torch.randn(b, t, h, w, dim, device=dev)  # canonical
            if temporal:
                xin = xc.permute(0, 2, 3, 1, 4).reshape(b * h * w, t, dim)
            else:
                xin = xc.reshape(b * t, h * w, dim)
            xin = xin.clone().requires_grad_()
            wr = sd["dsconv.weight"].clone().requires_grad_(); br = sd["dsconv.bias"].clone().requires_grad_()
            ref = O.peg({"dsconv.weight": wr, "dsconv.bias": br}, "", xin, (b, t, h, w)) + xin
            def to_canon(z):
                z = z.detach()
                return (z.reshape(b, h, w, t, dim).permute(0, 3, 1, 2, 4) if temporal else z.reshape(b, t, h, w, dim)).reshape(-1, dim)
            got = ops.peg_fwd(xc.reshape(-1, dim), w27, sd["dsconv.bias"], (b, t, h, w), temporal)
            report(f"peg_fwd {b,t,h,w,dim} temporal={temporal}", got, to_canon(ref), 1e-5)
            dyc = torch.randn(b, t, h, w, dim, device=dev)
            dy_in = dyc.permute(0, 2, 3, 1, 4).reshape(b * h * w, t, dim) if temporal else dyc.reshape(b * t, h * w, dim)
            ref.backward(dy_in)
            dx, dxb = ops.peg_bwd_data(dyc.reshape(-1, dim), w27, (b, t, h, w), temporal, want_bf16=True)
            report("peg_bwd_data", dx, to_canon(xin.grad), 1e-5)
            dw = torch.zeros(27, dim, device=dev); dbias = torch.zeros(dim, device=dev)
            ops.peg_bwd_weight(xc.reshape(-1, dim), dyc.reshape(-1, dim), dw, dbias, (b, t, h, w), temporal)
            report("peg_bwd_weight", dw, wr.grad.reshape(dim, 27).t(), 1e-4); report("peg_bwd_bias", dbias, br.grad, 1e-4)

def attn_ref(q, kv, grid, heads, temporal, qs, ks, bias):
    b, t, h, w = grid
    inner = heads * 32
    def seqs(z):
        z = z.float().reshape(b, t, h, w, -1)
        return z.permute(0, 2, 3, 1, 4).reshape(b * h * w, t, -1) if temporal else z.reshape(b * t, h * w, -1)
    qq, kk, vv = seqs(q), seqs(kv[:, :inner]), seqs(kv[:, inner:])
    S, n, _ = qq.shape
    qq, kk, vv = (z.reshape(S, n, heads, 32).permute(0, 2, 1, 3) for z in (qq, kk, vv))
    qq = F.normalize(qq, dim=-1) * qs; kk = F.normalize(kk, dim=-1) * ks
    sim = torch.einsum("bhid,bhjd->bhij", qq, kk) * 8
    if bias is not None: sim = sim + bias
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), vv).permute(0, 2, 1, 3).reshape(S, n, inner)
    out = out.reshape(b, h, w, t, inner).permute(0, 3, 1, 2, 4) if temporal else out.reshape(b, t, h, w, inner)
    return out.reshape(-1, inner)

def bias_tables(heads, h, w):
    tab = torch.randn(heads, (2 * h - 1) * (2 * w - 1), device=dev) * 2
    pos = torch.stack(torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")).reshape(2, -1).t().to(dev)
    rel = pos[:, None, :] - pos[None, :, :]
    idx = (rel[..., 0] + h - 1) * (2 * w - 1) + (rel[..., 1] + w - 1)
    full = tab[:, idx]  # heads, n, n
    return tab, full.max(dim=-1).values.contiguous(), full

def t_attn():
    for (b, t, h, w, heads) in [(2, 5, 4, 4, 2), (2, 6, 6, 6, 4), (1, 24, 24, 24, 8), (3, 24, 24, 24, 8)]:
        tokens = b * t * h * w; inner = heads * 32
        q = torch.randn(tokens, inner, device=dev).bfloat16(); kv = torch.randn(tokens, 2 * inner, device=dev).bfloat16()
        qs = 1 + 0.1 * torch.randn(32, device=dev); ks = 1 + 0.1 * torch.randn(32, device=dev)
        tab, rowmax, full = bias_tables(heads, h, w)
        for temporal in (False, True):
            o, lse = ops.attn_fwd(q, kv, (b, t, h, w), heads, temporal, qs, ks, tab, rowmax)
            torch.cuda.synchronize()
            ref = attn_ref(q, kv, (b, t, h, w), heads, temporal, qs, ks, None if temporal else full)
            report(f"attn_fwd {b,t,h,w,heads} temporal={temporal}", o, ref, 2e-2)
            # backward
            qf = q.float().requires_grad_(); kvf = kv.float().requires_grad_()
            qsr = qs.clone().requires_grad_(); ksr = ks.clone().requires_grad_()
            fullr = None if temporal else full.clone().requires_grad_()
            refo = attn_ref(qf, kvf, (b, t, h, w), heads, temporal, qsr, ksr, fullr)
            d_o = torch.randn_like(refo).bfloat16()
            refo.backward(d_o.float())
            dqs = torch.zeros(32, device=dev); dks = torch.zeros(32, device=dev); dtab = torch.zeros_like(tab)
            dq, dkv = ops.attn_bwd(q, kv, o, lse, d_o, (b, t, h, w), heads, temporal, qs, ks, dqs, dks, tab, rowmax, dtab)
            torch.cuda.synchronize()
            report("  attn_bwd dq", dq, qf.grad, 3e-2); report("  attn_bwd dkv", dkv, kvf.grad, 3e-2)
            report("  attn_bwd dq_scale", dqs, qsr.grad, 2e-2); report("  attn_bwd dk_scale", dks, ksr.grad, 2e-2)
            if not temporal:
                pos = torch.stack(torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")).reshape(2, -1).t().to(dev)
                rel = pos[:, None, :] - pos[None, :, :]
                idx = ((rel[..., 0] + h - 1) * (2 * w - 1) + (rel[..., 1] + w - 1)).reshape(-1)
                ref_dtab = torch.zeros_like(tab).index_add_(1, idx, fullr.grad.reshape(heads, -1))
                report("  attn_bwd dbias_table", dtab, ref_dtab, 2e-2)

for fn in (t_layernorm, t_geglu, t_peg, t_attn):
    guard(fn)
print("ALL_OK" if allok else "SOME_FAILED")
sys.exit(0 if allok else 1)
