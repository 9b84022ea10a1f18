"""Parameter containers mirroring CTPA_CLIP/ct_clip/attention.py (same class names, constructor arguments and
state_dict keys). They hold weights only: the arithmetic of PEG / Attention / FeedForward / LayerNorm on the CT-CLIP
path runs in libctclip_sm100.so through ctpa_clip_b200.engine; calling a block on its own dispatches there too.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


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


class GEGLU(nn.Module):
    def forward(self, x):  # only reachable through FeedForward, which the engine evaluates as a whole
        raise RuntimeError("GEGLU is fused into the feed-forward kernels; call the FeedForward block")


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
        """full (heads, h*w, h*w) bias, as the reference returns it (gathered from the table)"""
        h, w = dimensions
        device = self.net[0][0].weight.device
        tab = self.table(h, w, device)
        return tab[:, pair_index(h, w, device)]


def pair_index(h, w, device):
    pos = torch.stack(torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")).reshape(2, -1).t()
    rel = pos[:, None, :] - pos[None, :, :]
    return (rel[..., 0] + h - 1) * (2 * w - 1) + (rel[..., 1] + w - 1)


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
        raise RuntimeError("Transformer blocks are evaluated by CTViT.encode on canonical (b,t,h,w,d) tokens; "
                           "call CTViT.encode / CTViT.forward")
