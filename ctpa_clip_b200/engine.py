"""Functional forward / backward of the CTViT encoder on the sm_100a kernels (libctclip_sm100.so).

Tokens stay in one canonical HBM layout [(b t h w), d] for the whole encoder: fp32 residual stream, bf16 tensor-core
operands. The spatial / temporal re-arrangements of the reference (ctvit.py:315,321,325,329) never materialise — the
PEG and attention kernels do the index arithmetic. Every function cites the reference lines it replaces.

`encoder_forward` returns the output plus a `ctx` object with the saved activations; `encoder_backward` consumes it
and returns gradients keyed by the reference's parameter names. `EncodeFunction` wraps both for torch.autograd.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .shadow import bf16_of


# GEGLU inside the FF1 GEMM's epilogue / the FF2 dgrad GEMM's epilogue (CTCLIP_FUSE_GEGLU=0: separate row-wise kernels, A/B)
FUSE_GEGLU = os.environ.get("CTCLIP_FUSE_GEGLU", "1") != "0"


def ff_pad(n: int) -> int:
    return (n + 7) // 8 * 8


class LayerWeights:
    """bf16 / re-laid-out operand copies of one transformer layer (derived caches, never saved)."""

    def __init__(self, peg, attn, ff, dim, keep=None):
        """keep: dict owned by the caller that outlives this object — holds the zero-padded operand buffers, so that a
        rebuild after an optimiser step only re-copies the payload (and nothing at all for the plain casts, which are views
        of the optimiser's bf16 mirror when there is one)"""
        inner2, _ = ff[1].weight.shape
        self.ffi = inner2 // 2
        self.ffp = ff_pad(self.ffi)
        dev = ff[1].weight.device
        keep = keep if keep is not None else {}
        self.w27 = peg.dsconv.weight.detach().reshape(dim, 27).t().contiguous()           # [27, dim] fp32
        self.peg_bias = peg.dsconv.bias.detach()
        self.gamma = attn.norm.gamma.detach()
        self.wq = bf16_of(attn.to_q.weight)                                              # [inner, dim]
        self.wkv = bf16_of(attn.to_kv.weight)                                            # [2 inner, dim]
        self.wout = bf16_of(attn.to_out.weight)                                          # [dim, inner]
        self.q_scale = attn.q_scale.detach()
        self.k_scale = attn.k_scale.detach()
        self.ff_g = ff[0].weight.detach()
        self.ff_b = ff[0].bias.detach()
        w1 = bf16_of(ff[1].weight)
        if "w1p" not in keep:
            keep["w1p"] = torch.zeros(2 * self.ffp, dim, device=dev, dtype=torch.bfloat16)  # [x rows | gate rows], zero padded
            keep["w2p"] = torch.zeros(dim, self.ffp, device=dev, dtype=torch.bfloat16)
        w1p, w2p = keep["w1p"], keep["w2p"]
        # payload copies are collected by the caller and done in ONE launch (ops.copy2d_batch)
        self.repack = [(w1[: self.ffi], w1p[: self.ffi]), (w1[self.ffi:], w1p[self.ffp: self.ffp + self.ffi]),
                       (bf16_of(ff[4].weight), w2p[:, : self.ffi])]
        self.w1p, self.w2p = w1p, w2p


class EncoderWeights:
    def __init__(self, vit):
        dim = vit.dim
        pe = vit.to_patch_emb
        self.dim = dim
        self.heads = vit.heads
        self.pt, self.ps = vit.temporal_patch_size, vit.patch_size[0]
        self.pdim = pe[1].weight.numel()
        self.pe_g, self.pe_b = pe[1].weight.detach(), pe[1].bias.detach()
        kp = ff_pad(self.pdim)
        keep = vit.__dict__.setdefault("_operand_buffers", {})
        if kp == self.pdim:
            self.pe_w = bf16_of(pe[2].weight)
        else:
            if "pe_w" not in keep:
                keep["pe_w"] = torch.zeros(dim, kp, device=pe[2].weight.device, dtype=torch.bfloat16)
            keep["pe_w"][:, : self.pdim] = bf16_of(pe[2].weight)
            self.pe_w = keep["pe_w"]
        self.pe_bias = pe[2].bias.detach()
        self.pe_g2, self.pe_b2 = pe[3].weight.detach(), pe[3].bias.detach()
        self.spatial = [LayerWeights(l[0], l[1], l[3], dim, keep.setdefault(("s", i), {}))
                        for i, l in enumerate(vit.enc_spatial_transformer.layers)]
        self.temporal = [LayerWeights(l[0], l[1], l[3], dim, keep.setdefault(("t", i), {}))
                         for i, l in enumerate(vit.enc_temporal_transformer.layers)]
        pairs = [pr for L in self.spatial + self.temporal for pr in L.repack]
        self._repack_sources = [s for s, _ in pairs]     # keep fresh casts alive until the copy has run
        ops.copy2d_batch(pairs, keep)
        self.s_out = vit.enc_spatial_transformer.norm_out.gamma.detach()
        self.t_out = vit.enc_temporal_transformer.norm_out.gamma.detach()
        embed = vit.vq._codebook.embed.detach()[0]
        self.embed = embed
        self.embed_n, self.embed_nb, _ = ops.l2norm_rows(embed.contiguous(), want_f32=True, want_bf16=True)


class LayerCtx:
    __slots__ = ("x", "x1", "x2", "xn", "xr", "q", "kv", "o", "lse", "xf", "h1", "u")


class EncoderCtx:
    pass


# ------------------------------------------------------------------------------------------------ forward
def layer_forward(x, L: LayerWeights, grid, heads, temporal, tab, rowmax, save: bool, attn_grid=None):
    """x = peg(x)+x ; x = attn(x)+x ; x = ff(x)+x   (attention.py:322-331) on canonical fp32 tokens [T, dim].
    attn_grid (block-level `Transformer.forward` only): the attention sequences as a separate (b, t, h, w) grid when the
    caller's memory layout is not the canonical one (the PEG then reinterprets the flat buffer as `grid`, attention.py:69-70)"""
    x1 = ops.peg_fwd(x, L.w27, L.peg_bias, grid, temporal)                                   # attention.py:324
    if attn_grid is not None:
        grid = attn_grid
    xn, xr, _ = ops.layernorm_fwd(x1, L.gamma, None, want_bf16=True, want_raw_bf16=True)     # :141 (q side), :139 (raw kv)
    q = ops.gemm(xn, L.wq)                                                                   # :143 to_q
    kv = ops.gemm(xr, L.wkv)                                                                 # :143 to_kv
    o, lse = ops.attn_fwd(q, kv, grid, heads, temporal, L.q_scale, L.k_scale, tab, rowmax)   # :145-180
    x2 = ops.gemm(o, L.wout, out_dtype=torch.float32, resid=x1)                              # :181 + residual :326
    xf, _, _ = ops.layernorm_fwd(x2, L.ff_g, L.ff_b)                                         # :47
    if FUSE_GEGLU:
        h1, u = ops.gemm_geglu(xf, L.w1p)                                                    # :48 with GEGLU (:39-42) in the epilogue
    else:
        h1 = ops.gemm(xf, L.w1p)                                                             # :48
        u = ops.geglu_fwd(h1)                                                                # :39-42
    x3 = ops.gemm(u, L.w2p, out_dtype=torch.float32, resid=x2)                               # :51 + residual :331
    c = None
    if save:
        c = LayerCtx()
        c.x, c.x1, c.x2, c.xn, c.xr, c.q, c.kv, c.o, c.lse, c.xf, c.h1, c.u = x, x1, x2, xn, xr, q, kv, o, lse, xf, h1, u
    return x3, c


def bias_tables(vit, h, w, device=None):
    """(2h-1)(2w-1) relative-position bias table per head, its per-query-position maximum over keys, and the MLP
    activations kept for the backward (attention.py:257-276 on the distinct offsets; `ctclip_cpb_table_fwd`)"""
    return vit.spatial_rel_pos_bias.table_fwd(h, w)


def patch_embed_forward(W: EncoderWeights, video, save: bool):
    """to_patch_emb: patchify + LN + Linear + LN (ctvit.py:169-174) -> fp32 tokens [T, dim]"""
    a = ops.patch_ln_fwd(video, W.pe_g, W.pe_b, W.pt, W.ps)
    y0 = ops.gemm(a, W.pe_w, out_dtype=torch.float32, bias=W.pe_bias)
    _, _, x = ops.layernorm_fwd(y0, W.pe_g2, W.pe_b2, want_bf16=False, want_f32=True)
    return x, (a, y0) if save else None


def encoder_forward(vit, W: EncoderWeights, video, save: bool, tab=None, rowmax=None):
    """CTViT.to_patch_emb + CTViT.encode (ctvit.py:409,417 / 306-331). Returns (tokens fp32 [T, dim], ctx)."""
    b, _, f, hh, ww = video.shape
    t, h, w = f // W.pt, hh // W.ps, ww // W.ps
    grid = (b, t, h, w)
    ctx = EncoderCtx()
    ctx.grid, ctx.video = grid, video
    x, ctx.pe = patch_embed_forward(W, video, save)
    if tab is None:
        with torch.no_grad():
            tab, rowmax, cpb_acts = bias_tables(vit, h, w, video.device)
    else:
        cpb_acts = None
    ctx.tab, ctx.rowmax, ctx.cpb_acts = tab, rowmax, cpb_acts
    ctx.spatial, ctx.temporal = [], []
    for L in W.spatial:
        x, c = layer_forward(x, L, grid, W.heads, False, tab, rowmax, save)
        ctx.spatial.append(c)
    ctx.s_pre = x if save else None
    _, _, x = ops.layernorm_fwd(x, W.s_out, None, want_bf16=False, want_f32=True)          # norm_out, attention.py:333
    for L in W.temporal:
        x, c = layer_forward(x, L, grid, W.heads, True, None, None, save)
        ctx.temporal.append(c)
    ctx.t_pre = x if save else None
    _, _, x = ops.layernorm_fwd(x, W.t_out, None, want_bf16=False, want_f32=True)
    return x, ctx


def vq_forward(W: EncoderWeights, tokens, stats=None):
    """cosine-sim nearest code (ctvit.py:427): returns (indices int32 [T], inv_norm [T])"""
    _, xb, inv = ops.l2norm_rows(tokens, want_bf16=True)
    top2, tile_n = ops.gemm_top2(xb, W.embed_nb)
    idx = ops.vq_finalize(top2, tile_n, tokens, W.embed_n, stats=stats)
    return idx, inv


# ------------------------------------------------------------------------------------------------ backward
class GradStore:
    """fp32 gradient accumulators. Default: fresh zero tensors in kernel layout, converted to reference parameter shapes
    at the end and handed to autograd. Direct mode (CTClipTrainStep): when the kernel layout IS the parameter layout the
    accumulator is the parameter's existing .grad (flat arena, zeroed by the optimiser kernel) and autograd gets None."""

    def __init__(self, device, params=None, direct=False, ready=None):
        self.device = device
        self.g = {}
        self.params = params or {}
        self.direct = direct
        self.ready = ready if direct else None    # callable(list of parameters): their .grad in the arena is final

    def direct_grad(self, name, shape=None):
        """the parameter's own .grad if direct accumulation is possible for `name` (optionally: in this shape)"""
        if not self.direct or name not in self.params:
            return None
        g = self.params[name].grad
        if g is None or not g.is_contiguous() or g.dtype != torch.float32:
            return None
        if shape is not None and tuple(g.shape) != tuple(shape):
            return None
        return g

    def fresh(self, name, shape):
        """always a new zero tensor (for gradients that are also an INPUT of a later kernel of this backward)"""
        t = torch.zeros(shape, device=self.device, dtype=torch.float32)
        self.g[name] = t
        return t

    def zeros(self, name, shape):
        g = self.direct_grad(name, shape)
        if g is not None:
            return g
        t = torch.zeros(shape, device=self.device, dtype=torch.float32)
        self.g[name] = t
        return t


LAYER_PARAM_NAMES = ("0.dsconv.bias", "1.q_scale", "1.k_scale", "1.norm.gamma", "1.to_q.weight", "1.to_kv.weight",
                     "1.to_out.weight", "3.0.weight", "3.0.bias", "3.1.weight", "3.4.weight")


def layer_backward(g, g_bf, c: LayerCtx, L: LayerWeights, grid, heads, temporal, tab, rowmax, dtab, gs: GradStore, prefix,
                   dim):
    """gradient of one [PEG, attention, feed-forward] layer; g = dL/dx3 fp32 [T, dim] (g_bf: its bf16 copy, written by
    the kernel that produced g); returns (dL/dx, its bf16 copy)"""
    inner = L.wq.shape[0]
    if g_bf is None:
        g_bf = ops.cast_bf16(g)
    # ---- feed-forward (attention.py:44-52)
    du = ops.gemm(g_bf, L.w2p, b_t=True)                                                     # [T, ffp]
    dw2 = gs.direct_grad(prefix + "3.4.weight")
    if dw2 is not None:   # un-padded [dim, ffi] parameter gradient: the GEMM simply stops at column ffi of u
        ops.gemm(g_bf, c.u[:, : L.ffi], a_t=True, b_t=True, out=dw2, accumulate=True, splits=0)
    else:
        dw2 = gs.zeros(prefix + "3.4.weight", (dim, L.ffp))
        ops.gemm(g_bf, c.u, a_t=True, b_t=True, out=dw2, accumulate=True, splits=0)
    dh1 = ops.geglu_bwd(c.h1, du)
    dxf = ops.gemm(dh1, L.w1p, b_t=True)                  # [T, dim] bf16: only LayerNorm's backward reads it
    # one GEMM over both padded halves [x rows | gate rows] (two half-size GEMMs into the parameter's own gradient would
    # read xf twice and fill the SMs worse); un-padded into the parameter's gradient right away in direct mode (so the layer's
    # whole gradient span is final when the layer is), else handed to autograd at the end
    dw1_direct = gs.direct_grad(prefix + "3.1.weight", (2 * L.ffi, dim))
    if dw1_direct is not None:
        dw1 = torch.zeros((2 * L.ffp, dim), device=g.device, dtype=torch.float32)
    else:
        dw1 = gs.fresh(prefix + "3.1.weight", (2 * L.ffp, dim))
    ops.gemm(dh1, c.xf, a_t=True, b_t=True, out=dw1, accumulate=True, splits=0)
    if dw1_direct is not None:
        dw1_direct[: L.ffi] += dw1[: L.ffi]
        dw1_direct[L.ffi:] += dw1[L.ffp: L.ffp + L.ffi]
    dg = gs.zeros(prefix + "3.0.weight", (dim,))
    db = gs.zeros(prefix + "3.0.bias", (dim,))
    g2, g2_bf = ops.layernorm_bwd(dxf, c.x2, L.ff_g, add_in=g, dgamma=dg, dbeta=db, want_bf16=True)
    del dxf, dh1, du
    # ---- attention (attention.py:127-181)
    d_o = ops.gemm(g2_bf, L.wout, b_t=True)                                                  # [T, inner]
    dwo = gs.zeros(prefix + "1.to_out.weight", (dim, inner))
    ops.gemm(g2_bf, c.o, a_t=True, b_t=True, out=dwo, accumulate=True, splits=0)
    dqs = gs.zeros(prefix + "1.q_scale", (32,))
    dks = gs.zeros(prefix + "1.k_scale", (32,))
    dq, dkv = ops.attn_bwd(c.q, c.kv, c.o, c.lse, d_o, grid, heads, temporal, L.q_scale, L.k_scale, dqs, dks, tab, rowmax,
                           dtab)
    dxn = ops.gemm(dq, L.wq, b_t=True)                    # bf16, consumed by the LayerNorm backward below
    dwq = gs.zeros(prefix + "1.to_q.weight", (inner, dim))
    ops.gemm(dq, c.xn, a_t=True, b_t=True, out=dwq, accumulate=True, splits=0)
    dwkv = gs.zeros(prefix + "1.to_kv.weight", (2 * inner, dim))
    ops.gemm(dkv, c.xr, a_t=True, b_t=True, out=dwkv, accumulate=True, splits=0)
    dgam = gs.zeros(prefix + "1.norm.gamma", (dim,))
    g1, _ = ops.layernorm_bwd(dxn, c.x1, L.gamma, add_in=g2, dgamma=dgam)
    # += d(raw kv input): in place, so the fire-and-forget vector REDs (one fp32 add per element, the same sum as the
    # residual epilogue) replace the epilogue's load - add - store round trip (157 -> 119 us per call, cold L2)
    ops.gemm(dkv, L.wkv, b_t=True, out=g1, accumulate=True)
    del dxn, dq, dkv, d_o, g2, g2_bf
    # ---- PEG (attention.py:63-84)
    dw27 = gs.zeros(prefix + "0.dsconv.weight", (27, dim))
    dpb = gs.zeros(prefix + "0.dsconv.bias", (dim,))
    ops.peg_bwd_weight(c.x, g1, dw27, dpb, grid, temporal)
    out = ops.peg_bwd_data(g1, L.w27, grid, temporal, want_bf16=True)
    if gs.ready is not None:   # every gradient this layer wrote straight into the arena is final: its all-reduce may start
        done = [gs.params[prefix + n] for n in LAYER_PARAM_NAMES if gs.direct_grad(prefix + n) is not None]
        if done:
            gs.ready(done)
    return out


def encoder_backward(vit, W: EncoderWeights, ctx: EncoderCtx, g_tokens):
    """backward of encoder_forward; g_tokens = dL/d(tokens) fp32 [T, dim]. Returns {reference param name: grad}."""
    dim = W.dim
    dev = g_tokens.device
    gs = GradStore(dev, params=dict(vit.named_parameters()), direct=getattr(vit, "direct_grad", False),
                   ready=getattr(vit, "grad_ready", None))
    grid = ctx.grid
    v = ""
    # temporal transformer
    dgo = gs.zeros("enc_temporal_transformer.norm_out.gamma", (dim,))
    g, g_bf = ops.layernorm_bwd(g_tokens, ctx.t_pre, W.t_out, dgamma=dgo, want_bf16=True)
    for i in reversed(range(len(W.temporal))):
        g, g_bf = layer_backward(g, g_bf, ctx.temporal[i], W.temporal[i], grid, W.heads, True, None, None, None, gs,
                                 f"enc_temporal_transformer.layers.{i}.", dim)
        ctx.temporal[i] = None
    dgo = gs.zeros("enc_spatial_transformer.norm_out.gamma", (dim,))
    g, g_bf = ops.layernorm_bwd(g, ctx.s_pre, W.s_out, dgamma=dgo, want_bf16=True)
    dtab = torch.zeros_like(ctx.tab)
    for i in reversed(range(len(W.spatial))):
        g, g_bf = layer_backward(g, g_bf, ctx.spatial[i], W.spatial[i], grid, W.heads, False, ctx.tab, ctx.rowmax, dtab, gs,
                                 f"enc_spatial_transformer.layers.{i}.", dim)
        ctx.spatial[i] = None
    # patch embedding (ctvit.py:169-174)
    a, y0 = ctx.pe
    dg2 = gs.zeros("to_patch_emb.3.weight", (dim,))
    db2 = gs.zeros("to_patch_emb.3.bias", (dim,))
    dy0, dy0_bf = ops.layernorm_bwd(g, y0, W.pe_g2, dgamma=dg2, dbeta=db2, want_bf16=True)
    dbias = gs.fresh("to_patch_emb.2.bias", (dim,))
    ops.colsum(dy0, dbias)
    dwp = gs.fresh("to_patch_emb.2.weight", (dim, a.shape[1]))   # this step's dW alone feeds patch_ln_param_grad below
    ops.gemm(dy0_bf, a, a_t=True, b_t=True, out=dwp, accumulate=True, splits=0)
    dg1 = gs.zeros("to_patch_emb.1.weight", (W.pdim,))
    db1 = gs.zeros("to_patch_emb.1.bias", (W.pdim,))
    dwp_c = dwp[:, : W.pdim].contiguous() if a.shape[1] != W.pdim else dwp
    ops.patch_ln_param_grad(vit.to_patch_emb[2].weight.detach().contiguous(), dwp_c, dbias, W.pe_g, W.pe_b, dg1, db1)
    # continuous position bias MLP on the (2h-1)(2w-1) distinct offsets (`ctclip_cpb_table_bwd`)
    _, _, h, w = grid
    acts = ctx.cpb_acts
    if acts is None:
        acts = vit.spatial_rel_pos_bias.table_fwd(h, w)[2]
    out = {}
    for name, gr in vit.spatial_rel_pos_bias.table_bwd(h, w, acts, dtab).items():
        out["spatial_rel_pos_bias." + name] = gr
    # kernel layouts -> reference parameter shapes
    for name, t in gs.g.items():
        if name.endswith("0.dsconv.weight"):
            out[name] = t.t().reshape(dim, 1, 3, 3, 3)
        elif name.endswith("3.1.weight"):
            ffp = t.shape[0] // 2
            L = W.spatial[0]
            out[name] = torch.cat((t[: L.ffi], t[ffp: ffp + L.ffi]), dim=0)
        elif name.endswith("3.4.weight"):
            out[name] = t[:, : W.spatial[0].ffi]
        elif name == "to_patch_emb.2.weight":
            out[name] = t[:, : W.pdim]
        else:
            out[name] = t
    return out


def layernorm_module_forward(x, gamma, beta):
    shape = x.shape
    _, _, y = ops.layernorm_fwd(x.reshape(-1, shape[-1]).contiguous().float(), gamma, beta, want_bf16=False, want_f32=True)
    return y.reshape(shape)
