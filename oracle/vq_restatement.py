"""TEST INFRASTRUCTURE — CPU restatement of `vector_quantize_pytorch==1.1.2` as CT-CLIP uses it.

PARITY UNPINNED: the library is a third-party dependency of the reference (requirements.txt:9; call sites
CTPA_CLIP/ct_clip/ctvit.py:17,187,299,427) that is neither vendored under /root/reference nor installed in
this image, and the reference has no test that touches it. What follows restates the published algorithm of
`VectorQuantize(dim, codebook_size, use_cosine_sim=True)` -> `CosineSimCodebook` (heads=1, codebook_dim=dim so
project_in/out are identities; decay=0.8, eps=1e-5, kmeans_init=False, threshold_ema_dead_code=0,
commitment_weight=1.0, sample_codebook_temp=0, sync_codebook=False, learnable_codebook=False) from memory of
the upstream source; parity is anchored on the reference's own call sites (what is passed in, what is read out).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def l2norm(t: torch.Tensor) -> torch.Tensor:
    return F.normalize(t, p=2, dim=-1)


def cosine_codebook_init(codebook_size: int, dim: int, generator: torch.Generator | None = None) -> torch.Tensor:
    """l2norm(kaiming_uniform(1, codebook_size, dim)) — upstream `uniform_init` followed by l2norm."""
    # kaiming_uniform_ on a (1, C, D) tensor: fan_in = C*D... upstream calls nn.init.kaiming_uniform_(t) with
    # default a=0 -> bound = sqrt(6 / fan_in) where fan_in = size(1) * receptive(D). The scale is irrelevant after
    # l2norm, so only the uniform(-b, b) shape matters.
    t = torch.empty(1, codebook_size, dim)
    fan_in = codebook_size * dim
    bound = (6.0 / fan_in) ** 0.5
    t.uniform_(-bound, bound, generator=generator)
    return l2norm(t)


class CosineSimCodebook(nn.Module):
    def __init__(self, dim: int, codebook_size: int, decay: float = 0.8, eps: float = 1e-5):
        super().__init__()
        self.decay = decay
        self.eps = eps
        self.codebook_size = codebook_size
        self.register_buffer("initted", torch.Tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed", cosine_codebook_init(codebook_size, dim))

    def forward(self, x: torch.Tensor):
        with torch.autocast(device_type=x.device.type, enabled=False):
            x = x.float()
            shape = x.shape  # (b, n, d)
            flatten = l2norm(x.reshape(1, -1, shape[-1]))
            embed_n = l2norm(self.embed)
            dist = torch.einsum("hnd,hcd->hnc", flatten, embed_n)
            embed_ind = dist.argmax(dim=-1)  # temperature 0 -> plain argmax
            quantize = F.embedding(embed_ind[0], self.embed[0]).reshape(shape)  # raw (un-normalised) stored embed
            if self.training:
                onehot = F.one_hot(embed_ind, self.codebook_size).type(x.dtype)
                bins = onehot.sum(dim=1)
                self.cluster_size.data.lerp_(bins, 1 - self.decay)
                zero_mask = bins == 0
                bins = bins.masked_fill(zero_mask, 1.0)
                embed_sum = torch.einsum("hnd,hnc->hcd", flatten, onehot)
                embed_normalized = l2norm(embed_sum / bins.unsqueeze(-1))
                embed_normalized = torch.where(zero_mask.unsqueeze(-1), embed_n, embed_normalized)
                self.embed.data.lerp_(embed_normalized, 1 - self.decay)
            return quantize, embed_ind.reshape(shape[:-1])


class VectorQuantize(nn.Module):
    """Same constructor / forward / attribute surface the reference touches (ctvit.py:187,299,427)."""

    def __init__(self, dim, codebook_size, use_cosine_sim=True, decay=0.8, eps=1e-5, commitment_weight=1.0, **_):
        super().__init__()
        assert use_cosine_sim, "CT-CLIP builds the quantiser with use_cosine_sim=True (ctvit.py:187)"
        self.commitment_weight = commitment_weight
        self._codebook = CosineSimCodebook(dim, codebook_size, decay=decay, eps=eps)

    @property
    def codebook(self):
        return self._codebook.embed[0]

    def forward(self, x, mask=None):
        quantize, embed_ind = self._codebook(x)
        if self.training:
            quantize = x + (quantize - x).detach()
        loss = torch.zeros(1, device=x.device, requires_grad=self.training)
        if self.training and self.commitment_weight > 0:
            loss = loss + F.mse_loss(quantize.detach(), x) * self.commitment_weight
        return quantize, embed_ind, loss
