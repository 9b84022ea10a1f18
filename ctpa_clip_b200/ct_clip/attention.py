"""Modules mirroring CTPA_CLIP/ct_clip/attention.py (same class names, constructor arguments, forward signatures and
state_dict keys). On the CT-CLIP path (CTViT / CTCLIP forward + backward) they are parameter containers: the engine
(ctpa_clip_b200.engine) evaluates whole layers on libctclip_sm100.so without materialising the reference's rearranges.
Called on their own — `vit.enc_spatial_transformer(tokens, attn_bias=..., video_shape=...)` as
ctpa_report/vqa_meditron.py:107 does, `PEG(x, shape)`, `Attention(x, attn_bias=...)` — `forward` dispatches into the same
kernels with the reference's semantics on the caller's memory layout (inference surface: no autograd through a lone block).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..shadow import bf16_of


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


class LayerNorm(nn.Module):
    """gamma-only LayerNorm, beta is a zero buffer (reference attention.py:28-35)"""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))

    def forward(self, x):
        from .. import engine
        return engine.layernorm_module_forward(x, self.gamma, None)


def _no_block_autograd(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError("autograd through a lone block is not implemented: differentiate through CTViT / CTCLIP "
                                  "(hand-written backward of the whole encoder), or call the block under torch.no_grad()")


def _flat_f32(x):
    return x.reshape(-1, x.shape[-1]).contiguous().float()


class GEGLU(nn.Module):
    """x, gate = chunk(2); gelu(gate) * x (reference attention.py:39-42) on `ctclip_geglu_fwd`"""

    def forward(self, x):
        _no_block_autograd(x)
        shape, half = x.shape, x.shape[-1] // 2
        hp = (half + 7) // 8 * 8                      # the kernel wants [x | gate] halves at a 16-byte pitch
        flat = x.reshape(-1, 2 * half)
        h = torch.zeros((flat.shape[0], 2 * hp), device=x.device, dtype=torch.bfloat16)
        h[:, :half] = flat[:, :half]
        h[:, hp: hp + half] = flat[:, half:]
        u = ops.geglu_fwd(h)[:, :half]
        return u.to(x.dtype).reshape(*shape[:-1], half)


def FeedForward(dim, mult=4, dropout=0.0):
    """LayerNorm -> Linear(dim, 2*inner) -> GEGLU -> Dropout -> Linear(inner, dim)  (reference attention.py:44-52)"""
    if dropout != 0.0:
        raise NotImplementedError("ff_dropout != 0 is not used by CT-CLIP and not implemented")
    inner_dim = int(mult * (2 / 3) * dim)
    return nn.Sequential(
        nn.LayerNorm(dim),
        nn.Linear(dim, inner_dim * 2, bias=False),
        GEGLU(),
        nn.Dropout(dropout),
        nn.Linear(inner_dim, dim, bias=False),
    )


class PEG(nn.Module):
    """depth-wise causal 3x3x3 conv position generator (reference attention.py:56-84)"""

    def __init__(self, dim, causal=False):
        super().__init__()
        if not causal:
            raise NotImplementedError("CTViT builds PEG with peg_causal=True (ctvit.py:183); non-causal is not implemented")
        self.causal = causal
        self.dsconv = nn.Conv3d(dim, dim, 3, groups=dim)

    def forward(self, x, shape=None):
        """x: (b, t, h, w, d), or (b', n, d) with shape=(b, t, h, w): the flat buffer is REINTERPRETED as (b, t, h, w, d)
        exactly like the reference's `x.reshape(*shape, -1)` (attention.py:69-70), zero-padded (1,1,1,1,2,0) and convolved
        depth-wise over (t, h, w), causal on the first grid axis. Returns the convolution alone (the caller adds x).
        `ctclip_peg_fwd` computes x + conv(x) + bias in one pass; the residual is subtracted again here (fp32, one rounding)."""
        _no_block_autograd(x)
        needs_shape = x.ndim == 3
        if needs_shape and shape is None:
            raise ValueError("PEG.forward: a (b, n, d) input needs shape=(b, t, h, w)")
        grid = tuple(shape) if needs_shape else tuple(x.shape[:-1])
        dim = x.shape[-1]
        if len(grid) != 4 or x.numel() != grid[0] * grid[1] * grid[2] * grid[3] * dim:
            raise ValueError(f"PEG.forward: input of shape {tuple(x.shape)} is not a (b, t, h, w, d) grid {grid}")
        xf = _flat_f32(x)
        w27 = self.dsconv.weight.detach().reshape(dim, 27).t().contiguous().float()
        y = ops.peg_fwd(xf, w27, self.dsconv.bias.detach().float(), grid, False)
        return (y - xf).to(x.dtype).reshape(x.shape)


class Attention(nn.Module):
    """cosine-sim attention weights (reference attention.py:88-125)"""

    def __init__(self, dim, dim_context=None, dim_head=64, heads=8, causal=False, num_null_kv=0, norm_context=True,
                 dropout=0.0, scale=8):
        super().__init__()
        if causal or num_null_kv != 0 or dropout != 0.0 or scale != 8:
            raise NotImplementedError("only the CTViT self-attention configuration (non-causal, no null kv, scale 8) is implemented")
        if dim_head != 32:
            raise NotImplementedError("the sm_100a attention kernels are specialised for dim_head=32 (pretrained_model.py:25)")
        self.heads = heads
        self.causal = causal
        self.scale = scale
        inner_dim = dim_head * heads
        dim_context = default(dim_context, dim)
        self.norm = LayerNorm(dim)
        self.context_norm = LayerNorm(dim_context) if norm_context else nn.Identity()
        self.num_null_kv = num_null_kv
        self.null_kv = nn.Parameter(torch.randn(heads, 2 * num_null_kv, dim_head))
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim_context, inner_dim * 2, bias=False)
        self.q_scale = nn.Parameter(torch.ones(dim_head))
        self.k_scale = nn.Parameter(torch.ones(dim_head))
        self.to_out = nn.Linear(inner_dim, dim, bias=False)

    def forward(self, x, mask=None, context=None, attn_bias=None):
        """x: (b', n, d) -> to_out(softmax(8 * l2norm(q) q_scale . l2norm(k) k_scale + attn_bias) v)   (attention.py:127-181),
        K/V from the un-normalised x, Q from LayerNorm(x). attn_bias: (heads, n, n) relative-position bias of an (h, w) grid
        with h*w = n (what ContinuousPositionBias.forward returns); it is folded back into its (2h-1)(2w-1) table — the form
        the kernels gather from — and rejected if it is not translation-invariant."""
        _no_block_autograd(x)
        if mask is not None or context is not None:
            raise NotImplementedError("attention masks / cross-attention context are not used by CTViT and not implemented")
        b, n, dim = x.shape
        tab = rowmax = None
        grid = (b, 1, 1, n)
        if attn_bias is not None:
            hw = getattr(attn_bias, "_grid_hw", None)
            if hw is None:
                r = int(round(n ** 0.5))
                if r * r != n:
                    raise NotImplementedError("attn_bias of a non-square grid: pass the tensor returned by "
                                              "ContinuousPositionBias.forward (it carries its (h, w))")
                hw = (r, r)
            tab, rowmax = table_from_full_bias(attn_bias.detach().float(), *hw)
            grid = (b, 1, hw[0], hw[1])
        xf = _flat_f32(x)
        xn, xr, _ = ops.layernorm_fwd(xf, self.norm.gamma.detach().float(), None, want_bf16=True, want_raw_bf16=True)
        q = ops.gemm(xn, bf16_of(self.to_q.weight))
        kv = ops.gemm(xr, bf16_of(self.to_kv.weight))
        o, _ = ops.attn_fwd(q, kv, grid, self.heads, False, self.q_scale.detach().float(), self.k_scale.detach().float(),
                            tab, rowmax)
        out = ops.gemm(o, bf16_of(self.to_out.weight), out_dtype=torch.float32)
        return out.to(x.dtype).reshape(b, n, dim)


def leaky_relu(p=0.1):
    return nn.LeakyReLU(p)


class ContinuousPositionBias(nn.Module):
    """2 -> dim -> dim -> heads MLP on the signed log-distance grid (reference attention.py:229-276).

    The reference evaluates the MLP on all (h*w)^2 query/key pairs; only (2h-1)(2w-1) relative offsets are distinct,
    so `table` evaluates those and the attention kernels gather from the table."""

    def __init__(self, *, dim, heads, num_dims=2, layers=2, log_dist=True, cache_rel_pos=False):
        super().__init__()
        self.num_dims = num_dims
        self.log_dist = log_dist
        self.net = nn.ModuleList([])
        self.net.append(nn.Sequential(nn.Linear(self.num_dims, dim), leaky_relu()))
        for _ in range(layers - 1):
            self.net.append(nn.Sequential(nn.Linear(dim, dim), leaky_relu()))
        self.net.append(nn.Linear(dim, heads))
        self.cache_rel_pos = cache_rel_pos

    def rel_offsets(self, h, w, device):
        dy = torch.arange(-(h - 1), h, device=device)
        dx = torch.arange(-(w - 1), w, device=device)
        rel = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), dim=-1).reshape(-1, 2)
        if self.log_dist:
            rel = torch.sign(rel) * torch.log(rel.abs() + 1)
        return rel.float()

    def _mlp(self):
        if len(self.net) != 3 or self.num_dims != 2:
            raise NotImplementedError("the CUDA path implements the CTViT configuration: 2 -> dim -> dim -> heads (layers=2)")
        l0, l1, l2 = self.net[0][0], self.net[1][0], self.net[2]
        return l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias

    def table_fwd(self, h, w):
        """(table [heads, (2h-1)(2w-1)], rowmax [heads, h*w], saved activations) from `ctclip_cpb_table_fwd`;
        entry (dy+h-1)*(2w-1) + (dx+w-1) for the offset (dy, dx) = query - key"""
        with torch.no_grad():
            return ops.cpb_table_fwd(h, w, *[p.detach().float() for p in self._mlp()], log_dist=self.log_dist)

    def table_bwd(self, h, w, acts, dtable):
        """parameter gradients {name: grad} of the MLP from the table gradient (`ctclip_cpb_table_bwd`)"""
        _, _, w1, _, w2, _ = self._mlp()
        grads = ops.cpb_table_bwd(h, w, w1.detach().float(), w2.detach().float(), acts, dtable)
        names = ("net.0.0.weight", "net.0.0.bias", "net.1.0.weight", "net.1.0.bias", "net.2.weight", "net.2.bias")
        return dict(zip(names, grads))

    def table(self, h, w, device=None):
        return self.table_fwd(h, w)[0]

    def forward(self, *dimensions, device=None):
        """full (heads, h*w, h*w) bias, as the reference returns it (gathered from the table); tagged with its grid so that
        Attention / Transformer.forward can fold it back into the table without guessing (h, w)"""
        h, w = dimensions
        device = self.net[0][0].weight.device
        tab = self.table(h, w, device)
        out = tab[:, pair_index(h, w, device)]
        out._grid_hw = (h, w)
        return out


def pair_index(h, w, device):
    pos = torch.stack(torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")).reshape(2, -1).t()
    rel = pos[:, None, :] - pos[None, :, :]
    return (rel[..., 0] + h - 1) * (2 * w - 1) + (rel[..., 1] + w - 1)


def table_from_full_bias(bias, h, w):
    """(heads, n, n) relative-position bias of an (h, w) grid -> (table [heads, (2h-1)(2w-1)], rowmax [heads, n]); raises if
    the bias is not a function of the (dy, dx) offset alone (the kernels gather from the table)"""
    heads, n, n2 = bias.shape
    if n != h * w or n2 != n:
        raise ValueError(f"attn_bias {tuple(bias.shape)} does not belong to a {h} x {w} grid")
    idx = pair_index(h, w, bias.device)
    table = torch.zeros((heads, (2 * h - 1) * (2 * w - 1)), device=bias.device, dtype=torch.float32)
    table[:, idx.reshape(-1)] = bias.reshape(heads, -1)
    if not torch.equal(table[:, idx], bias):
        raise NotImplementedError("attn_bias is not translation-invariant: only relative-position biases "
                                  "(ContinuousPositionBias) are implemented")
    return table.contiguous(), bias.max(dim=-1).values.contiguous()


class Transformer(nn.Module):
    """[PEG, Attention, None, FeedForward] x depth + norm_out (reference attention.py:280-333)"""

    def __init__(self, dim, *, depth, dim_context=None, causal=False, dim_head=64, heads=8, ff_mult=4, peg=False,
                 peg_causal=False, attn_num_null_kv=2, has_cross_attn=False, attn_dropout=0.0, ff_dropout=0.0):
        super().__init__()
        if has_cross_attn or causal or not peg:
            raise NotImplementedError("only the CTViT encoder configuration (peg=True, no cross attention) is implemented")
        self.dim, self.heads, self.dim_head = dim, heads, dim_head
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PEG(dim=dim, causal=peg_causal),
                Attention(dim=dim, dim_head=dim_head, heads=heads, causal=causal, dropout=attn_dropout),
                None,
                FeedForward(dim=dim, mult=ff_mult, dropout=ff_dropout),
            ]))
        self.norm_out = LayerNorm(dim)

    def forward(self, x, video_shape=None, attn_bias=None, context=None, self_attn_mask=None,
                cross_attn_context_mask=None):
        """x: (b', n, d) tokens, video_shape = (b, t, h, w) (attention.py:312-333): per layer x = peg(x) + x (the flat buffer
        reinterpreted as video_shape), x = attn(x, attn_bias) + x over the n tokens of each of the b' sequences, x = ff(x) + x;
        then norm_out. Works on the caller's layout — '(b t) (h w) d' with attn_bias for the spatial transformer,
        '(b h w) t d' for the temporal one (ctvit.py:315-329) — through `engine.layer_forward`."""
        from .. import engine
        _no_block_autograd(x)
        if context is not None or self_attn_mask is not None or cross_attn_context_mask is not None:
            raise NotImplementedError("cross-attention context / masks are not used by CTViT and not implemented")
        if video_shape is None:
            raise ValueError("Transformer.forward: video_shape=(b, t, h, w) is required (PEG)")
        bprime, n, dim = x.shape
        grid = tuple(int(v) for v in video_shape)
        if bprime * n != grid[0] * grid[1] * grid[2] * grid[3]:
            raise ValueError(f"Transformer.forward: {tuple(x.shape)} tokens do not fill video_shape {grid}")
        tab = rowmax = None
        attn_grid = (bprime, 1, 1, n)
        if attn_bias is not None:
            hw = getattr(attn_bias, "_grid_hw", None) or (grid[2], grid[3])
            tab, rowmax = table_from_full_bias(attn_bias.detach().float(), *hw)
            attn_grid = (bprime, 1, hw[0], hw[1])
        with torch.no_grad():
            keep = self.__dict__.setdefault("_operand_buffers", {})
            Ls = [engine.LayerWeights(l[0], l[1], l[3], dim, keep.setdefault(i, {})) for i, l in enumerate(self.layers)]
            pairs = [pr for L in Ls for pr in L.repack]
            ops.copy2d_batch(pairs, keep)
            xf = _flat_f32(x)
            for L in Ls:
                xf, _ = engine.layer_forward(xf, L, grid, self.heads, False, tab, rowmax, False, attn_grid=attn_grid)
            _, _, xf = ops.layernorm_fwd(xf, self.norm_out.gamma.detach().float(), None, want_bf16=False, want_f32=True)
        return xf.to(x.dtype).reshape(bprime, n, dim)
