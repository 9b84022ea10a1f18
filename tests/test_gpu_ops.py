"""GPU (-m gpu): every kernel of libctclip_sm100.so, called through the C-ABI wrappers, against the CPU oracle /
fp32 torch restatements on the same seeded inputs. Tolerances: fp32 kernels 1e-4..1e-5 relative; kernels with bf16
operands 2-3e-2 of the tensor's max magnitude (bf16 has 8 mantissa bits: 2^-9 relative per rounding)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ctclip_oracle as O  # noqa: E402
from oracle import resample_oracle as R  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from ctpa_clip_b200 import ops as o
    return o


def close(got, ref, tol):
    err = (got.float().cpu() - ref.float().cpu()).abs().max().item()
    scale = max(ref.float().abs().max().item(), 1e-6)
    assert err <= tol * scale, f"err {err:.3e} vs scale {scale:.3e} (tol {tol})"


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("a_t", [False, True])
@pytest.mark.parametrize("b_t", [False, True])
@pytest.mark.parametrize("shape", [(128, 128, 64), (296, 520, 200), (1000, 96, 72), (8, 512, 4096)])
def test_gemm_layouts(ops, a_t, b_t, shape):
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = A.float() @ B.float().t()
    got = ops.gemm(A.t().contiguous() if a_t else A, B.t().contiguous() if b_t else B, a_t=a_t, b_t=b_t,
                   out_dtype=torch.float32)
    close(got, ref, 2e-5)


def test_gemm_epilogues_and_split_k(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 1500, 512, 1368
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    ref = A.float() @ B.float().t()
    close(ops.gemm(A, B, out_dtype=torch.float32, bias=bias, resid=resid), ref + bias + resid, 2e-5)
    close(ops.gemm(A, B), ref, 1e-2)                                   # bf16 store
    x = resid.clone()
    ops.gemm(A, B, out=x, resid=x)                                     # in-place residual stream update
    close(x, ref + resid, 2e-5)
    # wgrad shape: reduction over the long token axis, split-K with fp32 atomics
    T = 20000
    dy = torch.randn(T, 256, device="cuda", generator=g).bfloat16()
    xx = torch.randn(T, 512, device="cuda", generator=g).bfloat16()
    out = torch.ones(256, 512, device="cuda")
    ops.gemm(dy, xx, a_t=True, b_t=True, out=out, accumulate=True, splits=0)
    close(out, dy.float().t() @ xx.float() + 1, 1e-4)


@pytest.mark.parametrize("a_t", [False, True])
@pytest.mark.parametrize("b_t", [False, True])
@pytest.mark.parametrize("shape", [(3104, 1544, 576), (3208, 1304, 640)])
def test_gemm_cta_pair_layouts(ops, a_t, b_t, shape):
    """shapes that take the tcgen05.mma.cta_group::2 path (K >= 512, >= 148 tiles): odd number of m tiles (the second CTA
    of the last pair owns a phantom tile), ragged last m tile (8 valid rows) and ragged n tile; every operand layout;
    fp32 + bias + residual, bf16 and split-K atomic epilogues"""
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    ref = A.float() @ B.float().t()
    a_op, b_op = (A.t().contiguous() if a_t else A), (B.t().contiguous() if b_t else B)
    close(ops.gemm(a_op, b_op, a_t=a_t, b_t=b_t, out_dtype=torch.float32, bias=bias, resid=resid), ref + bias + resid, 2e-5)
    close(ops.gemm(a_op, b_op, a_t=a_t, b_t=b_t), ref, 1e-2)
    out = torch.ones(M, N, device="cuda")
    ops.gemm(a_op, b_op, a_t=a_t, b_t=b_t, out=out, accumulate=True, splits=3)
    close(out, ref + 1, 1e-4)


@pytest.mark.parametrize("shape", [(200, 176, 64), (128, 8, 64), (2000, 1368, 512), (19000, 1368, 512), (333, 520, 200)])
def test_gemm_geglu_forward_epilogue_is_bit_identical_to_the_unfused_pair(ops, shape):
    """FF1 with GEGLU in the epilogue (attention.py:39-48): h = [x | gate] equals the plain GEMM's output and u equals
    geglu_fwd(h) bit for bit, on non-pair tiles, CTA-pair tiles (K >= 512, >= 148 tiles), ragged M and a ragged last slab."""
    M, Nh, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + Nh + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(2 * Nh, K, device="cuda", generator=g) * 0.08).bfloat16()
    h_ref = ops.gemm(A, W)
    u_ref = ops.geglu_fwd(h_ref)
    h, u = ops.gemm_geglu(A, W)
    assert torch.equal(h, h_ref)
    assert torch.equal(u, u_ref)
    ref = A.float() @ W.float().t()
    close(u, ref[:, :Nh] * F.gelu(ref[:, Nh:]), 2e-2)


@pytest.mark.parametrize("dy_bf16", [True, False])
@pytest.mark.parametrize("with_add", [True, False])
def test_layernorm_bwd_ring_kernel_is_bit_identical(ops, dy_bf16, with_add, monkeypatch):
    """dim 512 runs on the cp.async-ring kernel (rows staged through shared memory two ahead); same arithmetic in the same
    order as the register-staged kernel, so dx is bit-identical; the parameter gradients are sums of the same per-row terms"""
    rows, dim = 5000, 512
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(rows, dim, device="cuda", generator=g) * 2 + 0.5
    dy = torch.randn(rows, dim, device="cuda", generator=g)
    dy = dy.bfloat16() if dy_bf16 else dy
    add = torch.randn(rows, dim, device="cuda", generator=g) if with_add else None
    gam = torch.randn(dim, device="cuda", generator=g)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("CTCLIP_LN_BWD_RING", flag)
        dg, db = torch.zeros(dim, device="cuda"), torch.zeros(dim, device="cuda")
        dx, dxb = ops.layernorm_bwd(dy, x, gam, add_in=add, dgamma=dg, dbeta=db, want_bf16=True)
        res.append((dx, dxb, dg, db))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    close(res[0][2], res[1][2], 1e-4); close(res[0][3], res[1][3], 1e-4)


@pytest.mark.parametrize("offset", [1024, 0])
def test_unpack12_restores_the_raw_scan_and_the_prep_output(ops, offset):
    """12-bit transfer format: ctclip_unpack12 of the host-packed bytes equals the clamped int16 scan bit for bit, and the
    data_prep output computed from it is bit-identical to the one computed from the original scan (the clamp lies outside
    the HU window)"""
    from ctpa_clip_b200.data_prep.pack12 import pack12, unpack12
    from ctpa_clip_b200.data_prep import preprocess_volumes
    g = torch.Generator().manual_seed(5)
    raw = torch.randint(-1024 if offset else 0, 3072 if offset else 4096, (2, 32, 32, 24), generator=g, dtype=torch.int16)
    raw.view(-1)[:4] = torch.tensor([-5000, 9000, -1024 if offset else 0, 3071 if offset else 4095], dtype=torch.int16)
    packed = pack12(raw, offset).cuda()
    got = unpack12(packed, raw.shape, offset)
    want = raw.clamp(-offset, 4095 - offset)
    assert torch.equal(got.cpu(), want)
    if offset:
        a = preprocess_volumes(raw.cuda(), 1.0, 0.0, 0.703125, 1.125)
        b = preprocess_volumes(got, 1.0, 0.0, 0.703125, 1.125)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))


def test_gemm_rejects_bad_alignment(ops):
    from ctpa_clip_b200._lib import CtclipError
    A = torch.zeros(16, 12, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(CtclipError):
        ops.gemm(A, A)                                                 # row stride 12 is not 16-byte aligned


# ------------------------------------------------------------------------------------------------ row-wise
@pytest.mark.parametrize("rows,dim", [(1000, 512), (77, 64), (0, 128), (1, 1024)])
def test_layernorm_fwd_bwd(ops, rows, dim):
    g = torch.Generator(device="cuda").manual_seed(rows + dim)
    x = torch.randn(rows, dim, device="cuda", generator=g) * 2 + 0.5
    gam = 1 + 0.1 * torch.randn(dim, device="cuda", generator=g)
    bet = 0.1 * torch.randn(dim, device="cuda", generator=g)
    y, raw, yf = ops.layernorm_fwd(x, gam, bet, want_bf16=True, want_raw_bf16=True, want_f32=True)
    if rows == 0:
        return
    ref = F.layer_norm(x, (dim,), gam, bet)
    close(yf, ref, 1e-5); close(y, ref, 1e-2); close(raw, x, 1e-2)
    xr, gr, br = x.clone().requires_grad_(), gam.clone().requires_grad_(), bet.clone().requires_grad_()
    dy = torch.randn(rows, dim, device="cuda", generator=g)
    F.layer_norm(xr, (dim,), gr, br).backward(dy)
    add = torch.randn(rows, dim, device="cuda", generator=g)
    dg, db = torch.zeros(dim, device="cuda"), torch.zeros(dim, device="cuda")
    dx, dxb = ops.layernorm_bwd(dy, x, gam, add_in=add, dgamma=dg, dbeta=db, want_bf16=True)
    close(dx, xr.grad + add, 1e-4); close(dxb, xr.grad + add, 1e-2); close(dg, gr.grad, 1e-4); close(db, br.grad, 1e-4)
    # upstream gradient handed over in bf16 (as the dgrad GEMM epilogue writes it): exact for the bf16-rounded dy
    dyb = dy.bfloat16()
    xr2, gr2, br2 = x.clone().requires_grad_(), gam.clone().requires_grad_(), bet.clone().requires_grad_()
    F.layer_norm(xr2, (dim,), gr2, br2).backward(dyb.float())
    dg2, db2 = torch.zeros(dim, device="cuda"), torch.zeros(dim, device="cuda")
    dx2, _ = ops.layernorm_bwd(dyb, x, gam, dgamma=dg2, dbeta=db2)
    close(dx2, xr2.grad, 1e-4); close(dg2, gr2.grad, 1e-4); close(db2, br2.grad, 1e-4)


def test_constant_rows_give_beta(ops):
    """air / padding patches are constant (-1): LayerNorm output must be exactly beta"""
    x = torch.full((5, 512), -1.0, device="cuda")
    gam, bet = torch.rand(512, device="cuda") + 0.5, torch.randn(512, device="cuda")
    _, _, yf = ops.layernorm_fwd(x, gam, bet, want_bf16=False, want_f32=True)
    assert torch.equal(yf, bet.expand(5, 512))


def test_geglu(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    rows, half = 513, 1368
    h = torch.randn(rows, 2 * half, device="cuda", generator=g).bfloat16()
    hf = h.float().requires_grad_()
    ref = hf[:, :half] * F.gelu(hf[:, half:])                         # attention.py:41-42: x * gelu(gate), gate = 2nd half
    close(ops.geglu_fwd(h), ref, 1e-2)
    du = torch.randn(rows, half, device="cuda", generator=g).bfloat16()
    ref.backward(du.float())
    close(ops.geglu_bwd(h, du), hf.grad, 1e-2)


# ------------------------------------------------------------------------------------------------ PEG
@pytest.mark.parametrize("grid", [(2, 5, 4, 4, 64), (2, 6, 6, 6, 128), (1, 24, 24, 24, 512), (1, 1, 3, 2, 32)])
@pytest.mark.parametrize("temporal", [False, True])
def test_peg_fwd_bwd_vs_oracle(ops, grid, temporal):
    b, t, h, w, dim = grid
    g = torch.Generator(device="cuda").manual_seed(sum(grid))
    wt = torch.randn(dim, 1, 3, 3, 3, device="cuda", generator=g) / 5
    bs = torch.randn(dim, device="cuda", generator=g) * 0.1
    w27 = wt.reshape(dim, 27).t().contiguous()
    xc = torch.randn(b, t, h, w, dim, device="cuda", generator=g)
    perm = (lambda z: z.permute(0, 2, 3, 1, 4).reshape(b * h * w, t, dim)) if temporal else (lambda z: z.reshape(b * t, h * w, dim))
    unperm = (lambda z: z.reshape(b, h, w, t, dim).permute(0, 3, 1, 2, 4)) if temporal else (lambda z: z.reshape(b, t, h, w, dim))
    xin = perm(xc).clone().requires_grad_()
    wr, br = wt.clone().requires_grad_(), bs.clone().requires_grad_()
    ref = O.peg({"dsconv.weight": wr, "dsconv.bias": br}, "", xin, (b, t, h, w)) + xin   # reference reshape semantics
    close(ops.peg_fwd(xc.reshape(-1, dim), w27, bs, (b, t, h, w), temporal), unperm(ref.detach()).reshape(-1, dim), 1e-5)
    dyc = torch.randn(b, t, h, w, dim, device="cuda", generator=g)
    ref.backward(perm(dyc))
    dx, dxb = ops.peg_bwd_data(dyc.reshape(-1, dim), w27, (b, t, h, w), temporal, want_bf16=True)
    close(dx, unperm(xin.grad).reshape(-1, dim), 1e-5)
    dw, dbias = torch.zeros(27, dim, device="cuda"), torch.zeros(dim, device="cuda")
    ops.peg_bwd_weight(xc.reshape(-1, dim), dyc.reshape(-1, dim), dw, dbias, (b, t, h, w), temporal)
    close(dw, wr.grad.reshape(dim, 27).t(), 1e-4); close(dbias, br.grad, 1e-4)


# ------------------------------------------------------------------------------------------------ attention
def _attn_inputs(grid, heads, seed):
    b, t, h, w = grid
    tokens, inner = b * t * h * w, heads * 32
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(tokens, inner, device="cuda", generator=g).bfloat16()
    kv = torch.randn(tokens, 2 * inner, device="cuda", generator=g).bfloat16()
    qs = 1 + 0.1 * torch.randn(32, device="cuda", generator=g)
    ks = 1 + 0.1 * torch.randn(32, device="cuda", generator=g)
    tab = torch.randn(heads, (2 * h - 1) * (2 * w - 1), device="cuda", generator=g) * 2
    from ctpa_clip_b200.ct_clip.attention import pair_index
    idx = pair_index(h, w, "cuda")
    return q, kv, qs, ks, tab, idx


def _attn_ref(q, kv, grid, heads, temporal, qs, ks, bias):
    """oracle attention core (attention.py:145-180) on canonical tokens"""
    b, t, h, w = grid
    inner = heads * 32
    def seqs(z):
        z = z.float().reshape(b, t, h, w, -1)
        return z.permute(0, 2, 3, 1, 4).reshape(b * h * w, t, -1) if temporal else z.reshape(b * t, h * w, -1)
    qq, kk, vv = seqs(q), seqs(kv[:, :inner]), seqs(kv[:, inner:])
    S, n, _ = qq.shape
    qq, kk, vv = (z.reshape(S, n, heads, 32).permute(0, 2, 1, 3) for z in (qq, kk, vv))
    qq, kk = F.normalize(qq, dim=-1) * qs, F.normalize(kk, dim=-1) * ks
    sim = torch.einsum("bhid,bhjd->bhij", qq, kk) * 8
    if bias is not None:
        sim = sim + bias
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), vv).permute(0, 2, 1, 3).reshape(S, n, inner)
    out = out.reshape(b, h, w, t, inner).permute(0, 3, 1, 2, 4) if temporal else out.reshape(b, t, h, w, inner)
    return out.reshape(-1, inner)


@pytest.mark.parametrize("grid,heads", [((2, 5, 4, 4), 2), ((2, 6, 6, 6), 4), ((1, 24, 24, 24), 8), ((2, 3, 13, 11), 1),
                                        ((1, 12, 5, 5), 2), ((1, 30, 3, 4), 3), ((1, 40, 3, 3), 2)])
@pytest.mark.parametrize("temporal", [False, True])
def test_attention_fwd_bwd(ops, grid, heads, temporal):
    b, t, h, w = grid
    q, kv, qs, ks, tab, idx = _attn_inputs(grid, heads, sum(grid) + heads)
    full = tab[:, idx]
    rowmax = full.amax(dim=-1).contiguous()
    o, lse = ops.attn_fwd(q, kv, grid, heads, temporal, qs, ks, tab, rowmax)
    qf, kvf = q.float().requires_grad_(), kv.float().requires_grad_()
    qsr, ksr = qs.clone().requires_grad_(), ks.clone().requires_grad_()
    fullr = None if temporal else full.clone().requires_grad_()
    ref = _attn_ref(qf, kvf, grid, heads, temporal, qsr, ksr, fullr)
    close(o, ref.detach(), 2.5e-2)
    d_o = torch.randn(ref.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9)).bfloat16()
    ref.backward(d_o.float())
    dqs, dks, dtab = torch.zeros(32, device="cuda"), torch.zeros(32, device="cuda"), torch.zeros_like(tab)
    dq, dkv = ops.attn_bwd(q, kv, o, lse, d_o, grid, heads, temporal, qs, ks, dqs, dks, tab, rowmax, dtab)
    close(dq, qf.grad, 3e-2); close(dkv, kvf.grad, 3e-2); close(dqs, qsr.grad, 2e-2); close(dks, ksr.grad, 2e-2)
    if not temporal:
        ref_dtab = torch.zeros_like(tab).index_add_(1, idx.reshape(-1), fullr.grad.reshape(heads, -1))
        close(dtab, ref_dtab, 2e-2)


@pytest.mark.parametrize("t", [3, 12, 24, 31])
def test_short_sequence_attention_equals_tcgen05_path(ops, t, monkeypatch):
    """temporal attention (n <= 32, no bias) runs on the one-warp-per-(sequence, head) mma.sync kernels; the tcgen05 kernels
    (CTCLIP_ATTN_SHORT=0) compute the same function with the same bf16 operand rounding"""
    grid, heads = (2, t, 5, 3), 4
    q, kv, qs, ks, tab, idx = _attn_inputs(grid, heads, 40 + t)
    d_o = torch.randn(q.shape[0], heads * 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)).bfloat16()
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("CTCLIP_ATTN_SHORT", flag)
        o, lse = ops.attn_fwd(q, kv, grid, heads, True, qs, ks, None, None)
        dqs, dks = torch.zeros(32, device="cuda"), torch.zeros(32, device="cuda")
        dq, dkv = ops.attn_bwd(q, kv, o, lse, d_o, grid, heads, True, qs, ks, dqs, dks, None, None, None)
        res.append((o, lse, dq, dkv, dqs, dks))
    for a, b2, tol in zip(res[0], res[1], (1e-2, 1e-3, 1.5e-2, 1.5e-2, 1e-2, 1e-2)):
        close(a, b2, tol)


def test_attention_large_bias_spread_does_not_underflow(ops):
    """bias is an unbounded MLP output: the per-row offset (row max of the table) must keep exp2 in range"""
    grid, heads = (1, 2, 6, 6), 1
    q, kv, qs, ks, tab, idx = _attn_inputs(grid, heads, 5)
    tab = tab * 40                                                   # spread of ~ +-150
    full = tab[:, idx]
    o, _ = ops.attn_fwd(q, kv, grid, heads, False, qs, ks, tab, full.amax(dim=-1).contiguous())
    ref = _attn_ref(q, kv, grid, heads, False, qs, ks, full)
    assert torch.isfinite(o.float()).all()
    close(o, ref, 2.5e-2)


# ------------------------------------------------------------------------------------------------ VQ
def test_vq_argmax_exact_including_near_ties(ops):
    """indices equal the fp32 arg-max for random tokens AND for adversarial near-ties that bf16 cannot separate"""
    from ctpa_clip_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(11)
    C, D, T = 1024, 512, 4096
    embed = F.normalize(torch.randn(C, D, device="cuda", generator=g), dim=-1)
    x = torch.randn(T, D, device="cuda", generator=g)
    # adversarial rows: a token almost equidistant from two codes (cosine gap ~1e-6 .. 1e-4)
    for r in range(256):
        a, b2 = int(torch.randint(0, C, (1,))), int(torch.randint(0, C, (1,)))
        eps = 10 ** (-6 + 2 * (r % 3) / 2)
        x[r] = embed[a] * (1 + eps) + embed[b2] + 0.01 * torch.randn(D, device="cuda", generator=g)
    class W: pass
    W.embed = embed
    W.embed_n, W.embed_nb, _ = ops.l2norm_rows(embed, want_f32=True, want_bf16=True)
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    idx, _ = engine.vq_forward(W, x.contiguous(), stats=stats)
    scores = F.normalize(x, dim=-1).double() @ F.normalize(embed, dim=-1).double().t()   # fp64 ground truth
    best = scores.argmax(dim=-1)
    got_score = scores.gather(1, idx.long()[:, None])[:, 0]
    # identical index, or an exact-tie-level alternative (fp32 cannot order gaps below ~1e-7)
    assert ((idx.long() == best) | ((scores.max(dim=-1).values - got_score) < 2e-7)).all()
    assert (idx.long() == best).float().mean() > 0.999
    assert stats[0].item() >= T                                       # at least one exact re-score per token


def test_vq_ema_matches_oracle(ops):
    g = torch.Generator(device="cuda").manual_seed(12)
    C, D, T = 64, 32, 500
    embed = F.normalize(torch.randn(1, C, D, device="cuda", generator=g), dim=-1)
    cluster = torch.rand(1, C, device="cuda", generator=g)
    x = torch.randn(T, D, device="cuda", generator=g)
    idx = torch.randint(0, C - 5, (T,), device="cuda", generator=g).to(torch.int32)     # last codes stay empty
    e_ref, c_ref = O.vq_ema(embed.cpu(), cluster.cpu(), x.cpu(), idx.cpu().long())
    _, _, inv = ops.l2norm_rows(x)
    e, c = embed.clone(), cluster.clone()
    bins, esum = ops.vq_ema(e[0], c[0], x, inv, idx)
    ops.vq_ema_update(e[0], c[0], bins, esum)
    close(e, e_ref, 1e-5); close(c, c_ref, 1e-5)


# ------------------------------------------------------------------------------------------------ loss / optimiser
@pytest.mark.parametrize("B", [1, 5, 64])
def test_clip_loss_fwd_bwd(ops, B):
    g = torch.Generator().manual_seed(B)
    T = F.normalize(torch.randn(B, 512, generator=g), dim=-1).requires_grad_()
    I = F.normalize(torch.randn(B, 512, generator=g), dim=-1).requires_grad_()
    tau = torch.tensor(1.0, requires_grad=True)
    ref = O.clip_loss(T, I, tau)
    ref.backward()
    loss, dT, dI, dtau = ops.clip_loss(T.detach().cuda(), I.detach().cuda(), tau.detach().reshape(1).cuda(), 0, B)
    assert abs(float(loss) - float(ref)) < 1e-5
    close(dT, T.grad, 1e-4 if B > 1 else 1.0); close(dI, I.grad, 1e-4 if B > 1 else 1.0)
    assert abs(float(dtau) - float(tau.grad)) < 1e-5
    # local row range (data-parallel share): rows [2, 4) of the same global batch
    if B >= 5:
        _, dT2, dI2, dtau2 = ops.clip_loss(T.detach().cuda(), I.detach().cuda(), tau.detach().reshape(1).cuda(), 2, 2)
        close(dT2, T.grad[2:4], 1e-4); close(dI2, I.grad[2:4], 1e-4)


def test_adam_with_clipping_matches_torch(ops):
    g = torch.Generator().manual_seed(0)
    n = 4096 * 3
    p0, grads = torch.randn(n, generator=g), [torch.randn(n, generator=g) * s for s in (3.0, 0.001, 1.0)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1.25e-6, betas=(0.9, 0.99), eps=1e-8)
    p, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step, gr in enumerate(grads, 1):
        ref.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([ref], 0.5)
        opt.step()
        gg, ns = gr.clone().cuda(), torch.zeros(1, device="cuda")
        ops.sumsq(gg, ns)
        ops.adam_step(p, gg, m, v, shadow, 1.25e-6, 0.9, 0.99, 1e-8, step, norm_sq=ns, max_norm=0.5)
        assert float(gg.abs().max()) == 0.0                            # gradient zeroed in the same pass
    assert (p.cpu() - ref.detach()).abs().max().item() < 2e-9
    close(shadow, ref.detach(), 1e-2)


def test_sumsq_is_deterministic_and_matches_torch(ops):
    """the clip norm is a fixed-order two-level sum (no atomics): bit-identical from call to call — and therefore on every
    data-parallel rank, whose all-reduced gradients are bit-identical — and equal to torch's within fp32 round-off"""
    g = torch.randn(3_000_004, device="cuda")
    outs = []
    for _ in range(4):
        o = torch.zeros(1, device="cuda")
        ops.sumsq(g, o)
        outs.append(o.clone())
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    assert abs(float(outs[0]) - float((g.double() ** 2).sum())) < 1e-5 * float(outs[0])


# ------------------------------------------------------------------------------------------------ data_prep
def test_resample_bit_exact_vs_golden_and_oracle(ops):
    G = np.load("tests/golden/resample.npz")
    for case in range(4):
        x, cur, want = G[f"case{case}_in"], G[f"case{case}_cur"], G[f"case{case}_out"]
        shape = R.resize_shape(x.shape, tuple(cur), (1.5, 0.75, 0.75))
        got = ops.prep_resample(torch.from_numpy(x)[None].cuda(), shape, layout="dhw")[0].cpu().numpy()
        assert got.shape == want.shape and (got.view(np.int32) == want.view(np.int32)).all()
    raw = np.ascontiguousarray(G["hu_raw"])
    for j in range(3):
        slope, intercept = (float(v) for v in G[f"hu{j}_params"])
        want = R.preprocess_volume(raw, slope, intercept, 0.9, 2.0)
        from ctpa_clip_b200.data_prep import preprocess_volumes
        got = preprocess_volumes(torch.from_numpy(raw)[None].cuda(), slope, intercept, 0.9, 2.0)[0].cpu().numpy()
        assert got.shape == want.shape and (got.view(np.int32) == want.view(np.int32)).all()


def test_resample_ragged_and_crop_pad(ops):
    rng = np.random.default_rng(4)
    for shape, out_grid, target in [((7, 33, 19), (11, 20, 31), None), ((10, 30, 21), (10, 30, 21), (8, 24, 24)),
                                    ((5, 20, 30), (6, 22, 27), (8, 24, 24)), ((1, 1, 1), (3, 2, 2), None)]:
        x = rng.random(shape, dtype=np.float32) * 2 - 1
        want = R.trilinear(x, out_grid)
        if target is not None:
            want = R.crop_pad(want, target, -1.0)
        got = ops.prep_resample(torch.from_numpy(x)[None].cuda(), out_grid, layout="dhw", target=target)[0].cpu().numpy()
        assert got.shape == want.shape and (got.view(np.int32) == want.view(np.int32)).all()


def test_resample_full_size_volume_bit_exact_and_properties(ops):
    """BASELINE config 4 shape: raw (512,512,320) int16 -> (240,480,480); one volume against the C oracle, plus
    size-independent properties on a batch (constant volumes stay constant, batch entries are independent)."""
    rng = np.random.default_rng(2)
    raw = rng.integers(-1024, 3071, size=(512, 512, 320), dtype=np.int16)
    from ctpa_clip_b200.data_prep import preprocess_volumes
    dev = torch.from_numpy(raw)[None].cuda()
    got = preprocess_volumes(dev, 1.0, -1024.0 + 1024.0, 0.703125, 1.125)
    assert tuple(got.shape) == (1, 240, 480, 480)
    want = R.preprocess_volume(raw, 1.0, 0.0, 0.703125, 1.125)
    assert (got[0].cpu().numpy().view(np.int32) == want.view(np.int32)).all()
    const = torch.full((2, 64, 64, 40), 300, dtype=torch.int16, device="cuda")
    out = preprocess_volumes(const, 1.0, 0.0, 0.703125, 1.125)
    assert torch.all(out == 0.3) or torch.allclose(out, torch.full_like(out, 0.3), atol=6e-8)
    pair = torch.stack((dev[0, :64, :64, :40], dev[0, 64:128, :64, :40])).contiguous()
    both = preprocess_volumes(pair, 1.0, 0.0, 0.703125, 1.125)
    single = preprocess_volumes(pair[1:].contiguous(), 1.0, 0.0, 0.703125, 1.125)
    assert torch.equal(both[1], single[0])


@pytest.mark.parametrize("shape,spacing,target,hu", [
    ((40, 48, 64), (0.703125, 1.125), None, (1.0, 0.0)),          # depth 64 -> 48 (down), 40x48 -> 37x45
    ((33, 70, 16), (0.9, 2.0), None, (1.0, -1024.0)),             # depth up-sampled 16 -> 21, ragged H / W
    ((50, 50, 24), (0.75, 1.5), None, (2.0, -1000.0)),            # identity grid on every axis
    ((64, 64, 40), (0.703125, 1.125), (24, 70, 48), (1.0, 0.0)),  # centre crop (h) + pad(-1) (w) window
    ((30, 30, 128), (1.6, 4.0), None, (0.5, 12.5)),               # strong down-sampling on every axis, fractional HU
    ((200, 136, 320), (0.703125, 1.125), None, (1.0, -1024.0)),   # production ratios (512->480-like 15/16, 320->240), two column tiles
    ((96, 144, 160), (0.72, 1.2), (110, 90, 100), (1.0, 0.0)),    # ragged last column tile + crop / pad window
    ((48, 100, 200), (0.703125, 1.125), None, (1.0, -1024.0)),    # 200 -> 150 planes: ragged last depth brick (150 = 3 x 48 + 6)
    ((40, 64, 104), (0.703125, 1.125), (70, 30, 50), (1.0, 100.0)),  # 104 -> 78: partial warp runs + crop (d) / crop (h) / crop (w)
])
def test_resample_marching_fast_path_equals_generic_and_oracle(ops, shape, spacing, target, hu, monkeypatch):
    """the depth-marching int16 (H,W,N) kernels — v2 (cp.async row ring, 128-bit shared-memory traffic, packed int16 / fp32x2
    arithmetic; two / one output columns per lane, 2 / 3 CTAs per SM), v1 with two output columns per lane, v1 with one
    (CTCLIP_PREP_X2=0) — and the generic brick kernel must agree bit for bit with the C oracle"""
    from ctpa_clip_b200.data_prep.preprocess import resize_shape
    rng = np.random.default_rng(11)
    raw = rng.integers(-2000, 3000, size=shape, dtype=np.int16)
    want = R.preprocess_volume(raw, hu[0], hu[1], spacing[0], spacing[1])
    if target is not None:
        want = R.crop_pad(want, target, -1.0)
    H, W, N = shape
    grid = resize_shape((N, H, W), (spacing[1], spacing[0], spacing[0]), (1.5, 0.75, 0.75))
    dev = torch.from_numpy(raw)[None].cuda()
    # v2 = "<columns per lane><CTAs per SM><raw rows by TMA (1) or cp.async (0)>" or "0" (v1 kernels)
    for force, v2, x2 in ((False, "121", "1"), (False, "120", "1"), (False, "131", "1"), (False, "221", "1"), (False, "220", "1"),
                          (False, "0", "1"), (False, "0", "0"), (True, "121", "1")):
        monkeypatch.setenv("CTCLIP_PREP_V2", "0" if v2 == "0" else "1")
        monkeypatch.setenv("CTCLIP_PREP_V2_NC", v2[0])
        monkeypatch.setenv("CTCLIP_PREP_V2_OCC", v2[1:2] or "2")
        monkeypatch.setenv("CTCLIP_PREP_V2_TMA", v2[2:3] or "0")
        monkeypatch.setenv("CTCLIP_PREP_X2", x2)
        got = ops.prep_resample(dev, grid, hu=hu, layout="hwn", target=target, force_generic=force)[0].cpu().numpy()
        assert got.shape == want.shape
        assert (got.view(np.int32) == want.view(np.int32)).all(), f"force_generic={force} v2={v2} x2={x2}"


# ---- DataLoader conversions on the GPU ---------------------------------------------------------------
def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()


@pytest.mark.parametrize("i", [0, 1])
def test_inference_loader_volume_equals_reference_golden(ops, i):
    """data_inference.py:78-122 fused into the resample kernel; golden = the reference loader's own output"""
    from ctpa_clip_b200.data_prep import inference_loader_volume
    g = np.load("tests/golden/loaders.npz")
    arr = g[f"infer{i}_q"].astype(np.float32) / np.float32(1024)
    got = inference_loader_volume(arr).cpu().numpy()
    assert got.shape == (1, 240, 480, 480)
    assert _sha(got) == g[f"infer{i}_sha256"].tobytes()


@pytest.mark.parametrize("i", [0, 1, 2])
def test_training_loader_volume_equals_reference_golden(ops, i):
    """data.py:114-192 (affine -> resize_array -> clip / 1000 -> crop / pad(-1) -> permute) in one kernel, bit-exact"""
    from ctpa_clip_b200.data_prep import training_loader_volume
    g = np.load("tests/golden/loaders.npz")
    s, ic, xy, z = (float(v) for v in g[f"train{i}_params"])
    got = training_loader_volume(g[f"train{i}_in"], s, ic, xy, z).cpu().numpy()
    assert got.shape == (1, 240, 480, 480)
    assert _sha(got) == g[f"train{i}_sha256"].tobytes()


def test_training_loader_volume_full_size_equals_oracle(ops):
    """a CT-RATE sized float32 scan (512 x 512 x 160, spacing 0.82 / 2.0: crop in h / w, pad in depth) against the C oracle"""
    from ctpa_clip_b200.data_prep import training_loader_volume
    rng = np.random.default_rng(3)
    arr = (rng.random((512, 512, 160), dtype=np.float32) * 3000 - 1400).astype(np.float32)
    want = R.training_loader_volume(arr, 1.0, -8.25, 0.82, 2.0)
    got = training_loader_volume(arr, 1.0, -8.25, 0.82, 2.0).cpu().numpy()
    assert (got.view(np.int32) == want.view(np.int32)).all()


def test_loader_ops_reject_int16_input(ops):
    with pytest.raises(Exception):
        ops.prep_resample(torch.zeros(1, 8, 8, 8, dtype=torch.int16, device="cuda"), (8, 8, 8), hu=(1.0, 0.0), post_op="clip_div")


def test_resample_to_target_equals_oracle(ops):
    """a15: direct (C, D, H, W) -> (C, 240, 480, 480) trilinear resample of the report generator, two channels, mixed up / down"""
    from ctpa_clip_b200.data_prep import resample_to_target
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 150, 300, 512), dtype=np.float32)
    got = resample_to_target(torch.from_numpy(x)).cpu().numpy()
    assert got.shape == (2, 240, 480, 480)
    for c in range(2):
        want = R.trilinear(x[c], (240, 480, 480))
        assert (got[c].view(np.int32) == want.view(np.int32)).all()


@pytest.mark.parametrize("rows,dim", [(4096, 768), (4096, 3072), (1000, 512), (37, 20), (5, 6)])
def test_column_sums(ops, rows, dim):
    """bias-gradient column sums (fp32 and bf16 inputs, tiled 16-byte path and the scalar fallback), accumulated in place"""
    g = torch.Generator(device="cuda").manual_seed(rows + dim)
    x = torch.randn(rows, dim, device="cuda", generator=g)
    out = torch.ones(dim, device="cuda")
    ops.colsum(x, out)
    close(out, 1 + x.double().sum(0).float(), 2e-5)
    xb = x.bfloat16()
    outb = torch.zeros(dim, device="cuda")
    ops.colsum_bf16(xb, outb)
    close(outb, xb.double().sum(0).float(), 2e-5)
    if dim % 3 == 0:                                  # a column slice of a packed [rows, 3 * d] matrix (BERT q / k / v biases)
        d = dim // 3
        outs = torch.zeros(d, device="cuda")
        ops.colsum_bf16(xb[:, d:2 * d], outs, dim=d, ld=dim)
        close(outs, xb[:, d:2 * d].double().sum(0).float(), 2e-5)


# ------------------------------------------------------------------------------------------------ continuous position bias
@pytest.mark.parametrize("h,w,dim,heads", [(24, 24, 512, 8), (3, 5, 64, 4), (2, 2, 96, 3)])
def test_cpb_table_fwd_bwd(ops, h, w, dim, heads):
    """ctclip_cpb_table_{fwd,bwd} vs the reference MLP (attention.py:229-276) evaluated by torch on ALL (h*w)^2 pairs in fp32:
    the table gathered to (heads, n, n) must equal the reference bias, and the parameter gradients must match autograd."""
    from ctpa_clip_b200.ct_clip.attention import ContinuousPositionBias, pair_index
    torch.manual_seed(h * 100 + w)
    cpb = ContinuousPositionBias(dim=dim, heads=heads).cuda()
    # reference restatement, full pair grid exactly as attention.py:257-276 builds it
    pos = torch.stack(torch.meshgrid(torch.arange(h, device="cuda"), torch.arange(w, device="cuda"), indexing="ij"))
    grid = pos.reshape(2, -1).t()
    rel = (grid[:, None, :] - grid[None, :, :]).float()
    rel = torch.sign(rel) * torch.log(rel.abs() + 1)
    x = rel
    for layer in cpb.net:
        x = layer(x)
    ref = x.permute(2, 0, 1)                                            # 'i j h -> h i j'
    tab, rowmax, acts = cpb.table_fwd(h, w)
    idx = pair_index(h, w, "cuda")
    close(tab[:, idx], ref, 2e-5)
    close(rowmax, ref.amax(dim=-1), 2e-5)
    close(cpb(h, w), ref, 2e-5)                                         # module forward = gathered table
    g = torch.randn(heads, h * w, h * w, device="cuda")
    ref_grads = torch.autograd.grad(ref, list(cpb.parameters()), g)
    dtab = torch.zeros_like(tab).index_add_(1, idx.reshape(-1), g.reshape(heads, -1))
    got = cpb.table_bwd(h, w, acts, dtab)
    for (name, _), rg in zip(cpb.named_parameters(), ref_grads):
        close(got[name], rg, 2e-4)


# ------------------------------------------------------------------------------------------------ zero-shot scoring
@pytest.mark.parametrize("V,P,d", [(32, 18, 512), (5, 3, 64), (1, 1, 4)])
def test_zero_shot_scores_kernel(ops, V, P, d):
    """ctclip_zero_shot_scores vs the reference's per-pathology pattern (ctclip_inference.py:318-330): softmax over the
    (present, absent) logit pair, element 0."""
    g = torch.Generator(device="cuda").manual_seed(V * 7 + P)
    I = F.normalize(torch.randn(V, d, device="cuda", generator=g), dim=-1)
    T = F.normalize(torch.randn(2 * P, d, device="cuda", generator=g), dim=-1)
    tau = torch.tensor([2.3], device="cuda")
    prob, logits = ops.zero_shot_scores(I, T, tau, want_logits=True)
    ref_logits = (I.double() @ T.double().t() * tau.double().exp()).view(V, P, 2)
    assert torch.allclose(logits.double(), ref_logits, atol=1e-4)
    assert torch.allclose(prob.double(), ref_logits.softmax(dim=-1)[..., 0], atol=1e-5)


# ------------------------------------------------------------------------------------------------ factor-gather weight gradient
@pytest.mark.parametrize("K", [8, 16, 32, 64])
def test_factor_gather_wgrad_gemm(ops, K):
    """dW[512, 294912] += dL_all^T E_all at the global batches of 1 / 2 / 4 / 8 ranks (ct_clip.LinearFunction.backward,
    factor_gather): MN-major bf16 operands, in-place fp32 residual epilogue, against torch fp32 (operands are exact in bf16,
    accumulation is fp32 -> 1e-5 of the result's scale)."""
    g = torch.Generator(device="cuda").manual_seed(K)
    dy = torch.randn(K, 512, device="cuda", generator=g).bfloat16()
    x = torch.randn(K, 294912, device="cuda", generator=g).bfloat16()
    w = torch.randn(512, 294912, device="cuda", generator=g)
    ref = w + dy.float().t() @ x.float()
    ops.gemm(dy, x, a_t=True, b_t=True, out=w, resid=w)
    close(w, ref, 1e-5)
