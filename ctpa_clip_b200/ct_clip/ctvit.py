"""CTViT image encoder — same constructor, forward signature, attributes and state_dict keys as
CTPA_CLIP/ct_clip/ctvit.py:117-436, evaluated by the sm_100a kernels of libctclip_sm100.so.

Only the encode half that CT-CLIP executes (`return_encoded_tokens=True`, `return_only_codebook_ids=True`) is implemented;
the decoder / VGG / GAN reconstruction half (ctvit.py:189-224, 333-375, 438-546) is out of scope and raises.
"""
from __future__ import annotations

from pathlib import Path

import torch
from torch import nn

from .. import engine, ops
from .attention import ContinuousPositionBias, Transformer


def pair(val):
    ret = (val, val) if not isinstance(val, tuple) else val
    assert len(ret) == 2
    return ret


class PatchEmbed(nn.Sequential):
    """Rearrange -> LayerNorm -> Linear -> LayerNorm (ctvit.py:169-174); callable on (b, 1, f, h, w) like the reference
    (ctpa_report/model_components.py:51 uses it stand-alone). Index 0 is a parameter-less placeholder for the Rearrange."""

    def __init__(self, pdim, dim, owner):
        super().__init__(nn.Identity(), nn.LayerNorm(pdim), nn.Linear(pdim, dim), nn.LayerNorm(dim))
        object.__setattr__(self, "_owner", owner)

    def forward(self, video):
        vit = self._owner
        W = vit.weights()
        b, _, f, hh, ww = video.shape
        x, _ = engine.patch_embed_forward(W, video.contiguous().float(), save=False)
        return x.reshape(b, f // W.pt, hh // W.ps, ww // W.ps, vit.dim)


class _CosineSimCodebook(nn.Module):
    def __init__(self, dim, codebook_size):
        super().__init__()
        from torch.nn import functional as F
        embed = torch.empty(1, codebook_size, dim)
        nn.init.kaiming_uniform_(embed)
        self.register_buffer("initted", torch.Tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed", F.normalize(embed, dim=-1))


class VectorQuantize(nn.Module):
    """buffers / attribute surface of vector_quantize_pytorch.VectorQuantize(dim, codebook_size, use_cosine_sim=True)
    (ctvit.py:187,299); decay 0.8. The assignment and EMA arithmetic run in the vq kernels (csrc/vq.cu)."""

    def __init__(self, dim, codebook_size, use_cosine_sim=True, decay=0.8):
        super().__init__()
        assert use_cosine_sim
        self.decay = decay
        self._codebook = _CosineSimCodebook(dim, codebook_size)

    @property
    def codebook(self):
        return self._codebook.embed[0]


class EncodeFunction(torch.autograd.Function):
    """CTViT.forward(video, return_encoded_tokens=True) [+ mean over t] as one autograd node over the CUDA engine."""

    @staticmethod
    def forward(ctx, vit, video, mode, *params):
        W = vit.weights()
        need_grad = any(ctx.needs_input_grad)
        video = video.contiguous().float()
        tokens, ectx = engine.encoder_forward(vit, W, video, save=need_grad)
        idx, inv = engine.vq_forward(W, tokens, stats=vit.vq_stats)
        if vit.force_indices is not None:  # test hook: quantise with externally supplied codes (parity with the VQ arg-max taken out)
            idx = vit.force_indices.reshape(-1).to(device=idx.device, dtype=torch.int32).contiguous()
        b, t, h, w = ectx.grid
        vit.last_indices = idx.view(b, t, h, w)
        if mode == "pooled":                                            # ct_clip.py:724,740
            out, out_bf = ops.vq_gather_mean(W.embed, idx, b, t, h * w)
            vit.last_pooled_bf16 = out_bf
        elif mode == "tokens":                                          # ctvit.py:433-436
            out = ops.vq_gather(W.embed, idx).view(b, t, h, w, vit.dim)
        else:
            raise ValueError(mode)
        if vit.training:                                                # EMA codebook update inside forward, as upstream
            vit._ema_update(tokens, inv, idx)
        ctx.vit, ctx.W, ctx.ectx, ctx.mode = vit, W, ectx, mode
        ctx.names = vit._param_names
        return out

    @staticmethod
    def backward(ctx, g):
        vit, W, ectx = ctx.vit, ctx.W, ctx.ectx
        b, t, h, w = ectx.grid
        if ctx.mode == "pooled":
            g_tokens = ops.pool_bwd(g.contiguous().float(), b, t, h * w, vit.dim)   # straight-through + mean backward
        else:
            g_tokens = g.reshape(-1, vit.dim).contiguous().float()
        grads = engine.encoder_backward(vit, W, ectx, g_tokens)
        ctx.ectx = None
        out = []
        shapes = vit._param_shapes          # ONE walk of the module tree (the property re-walks it on every access: 9 ms per step)
        for name in ctx.names:
            gr = grads.get(name)
            out.append(gr.reshape(shapes[name]) if gr is not None else None)
        return (None, None, None, *out)


class CTViT(nn.Module):
    def __init__(self, *, dim, codebook_size, image_size, patch_size, temporal_patch_size, spatial_depth, temporal_depth,
                 discr_base_dim=16, dim_head=64, heads=8, channels=1, use_vgg_and_gan=True, vgg=None,
                 discr_attn_res_layers=(16,), use_hinge_loss=True, attn_dropout=0.0, ff_dropout=0.0):
        super().__init__()
        if channels != 1:
            raise NotImplementedError("CT volumes are single-channel (pretrained_model.py:17-27); channels != 1 is not implemented")
        if attn_dropout != 0.0 or ff_dropout != 0.0:
            raise NotImplementedError("dropout is 0 in CT-CLIP (ctvit.py:136-137 defaults)")
        self.dim, self.heads, self.dim_head = dim, heads, dim_head
        self.image_size = pair(image_size)
        self.patch_size = pair(patch_size)
        patch_height, patch_width = self.patch_size
        if patch_height != patch_width:
            raise NotImplementedError("square spatial patches only")
        self.temporal_patch_size = temporal_patch_size
        self.spatial_rel_pos_bias = ContinuousPositionBias(dim=dim, heads=heads)
        image_height, image_width = self.image_size
        assert (image_height % patch_height) == 0 and (image_width % patch_width) == 0
        # present in the reference's state_dict, never executed on the CLIP path (ctvit.py:162-167, 189-197)
        self.to_patch_emb_first_frame = nn.Sequential(
            nn.Identity(), nn.LayerNorm(channels * patch_width * patch_height),
            nn.Linear(channels * patch_width * patch_height, dim), nn.LayerNorm(dim))
        pdim = channels * patch_width * patch_height * temporal_patch_size
        self.to_patch_emb = PatchEmbed(pdim, dim, self)
        kw = dict(dim=dim, dim_head=dim_head, heads=heads, attn_dropout=attn_dropout, ff_dropout=ff_dropout, peg=True,
                  peg_causal=True)
        self.enc_spatial_transformer = Transformer(depth=spatial_depth, **kw)
        self.enc_temporal_transformer = Transformer(depth=temporal_depth, **kw)
        self.vq = VectorQuantize(dim=dim, codebook_size=codebook_size, use_cosine_sim=True)
        self.to_pixels_first_frame = nn.Sequential(nn.Linear(dim, channels * patch_width * patch_height), nn.Identity())
        self.to_pixels = nn.Sequential(nn.Linear(dim, pdim), nn.Identity())
        self.use_vgg_and_gan = False  # the GAN / perceptual half is out of scope; no vgg.* / discr.* keys are created
        self._wcache = None
        self._wkey = None
        self.vq_stats = None
        self.last_indices = None
        self.last_pooled_bf16 = None
        self.force_indices = None
        self.ema_reduce = None  # optional callable(bins, embed_sum) -> work handle | None: cross-rank sum (async all-reduce)
        self.grad_ready = None  # optional callable(list of parameters): their .grad is final (direct-gradient mode, per layer)
        self._ema_pending = None

    # ---------------------------------------------------------------- reference surface
    @property
    def patch_height_width(self):
        return self.image_size[0] // self.patch_size[0], self.image_size[1] // self.patch_size[1]

    @property
    def image_num_tokens(self):
        h, w = self.patch_height_width
        return h * w

    def load(self, path):
        """ctvit.py:292-296. A reference CTViT built with use_vgg_and_gan=True (pretrained_model.py) also carries the
        `vgg.*` / `discr.*` weights of the reconstruction / GAN half, which does not exist here: those keys are dropped,
        everything else must match exactly (strict)."""
        path = Path(path)
        assert path.exists()
        sd = torch.load(str(path))
        sd = {k: v for k, v in sd.items() if not (k.startswith("vgg.") or k.startswith("discr."))}
        self.load_state_dict(sd)

    def get_video_patch_shape(self, num_frames, include_first_frame=True):
        patch_frames = 0
        if include_first_frame:
            num_frames -= 1
            patch_frames += 1
        patch_frames += num_frames // self.temporal_patch_size
        return (patch_frames, *self.patch_height_width)

    # ---------------------------------------------------------------- derived operand caches
    @property
    def _param_names(self):
        return [n for n, _ in self.named_parameters()]

    @property
    def _param_shapes(self):
        return {n: p.shape for n, p in self.named_parameters()}

    def weights(self) -> engine.EncoderWeights:
        self.finish_ema()
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple(
            (b.data_ptr(), b._version) for b in self.buffers())
        if self._wcache is None or key != self._wkey:
            with torch.no_grad():
                self._wcache = engine.EncoderWeights(self)
            self._wkey = key
        return self._wcache

    def invalidate_weights(self):
        self._wcache = None

    def finish_ema(self):
        """apply a pending (cross-rank reduced) EMA codebook update; no-op when nothing is pending"""
        pending = self.__dict__.get("_ema_pending")
        if pending is None:
            return
        self._ema_pending = None
        work, bins, esum = pending
        if work is not None:
            work.wait()
        cb = self.vq._codebook
        ops.vq_ema_update(cb.embed[0], cb.cluster_size[0], bins, esum, self.vq.decay)
        self._wcache = None  # the codebook operands are derived caches

    def _ema_update(self, tokens, inv, idx):
        cb = self.vq._codebook
        self.finish_ema()
        bins, esum = ops.vq_ema(cb.embed[0], cb.cluster_size[0], tokens, inv, idx, self.vq.decay)
        if self.ema_reduce is not None:
            # data parallel: the statistics are summed across ranks ASYNCHRONOUSLY (the callable returns a work handle); the
            # codebook itself is only needed by the next forward, so the update is applied by finish_ema() — which the trainer
            # calls with the gradient reduction — instead of stalling this forward on a 16.8 MB all-reduce
            self._ema_pending = (self.ema_reduce(bins, esum), bins, esum)
            return
        ops.vq_ema_update(cb.embed[0], cb.cluster_size[0], bins, esum, self.vq.decay)
        self._wcache = None  # the codebook operands are derived caches

    # ---------------------------------------------------------------- compute
    def encode(self, tokens):
        """spatial then temporal transformer on (b, t, h, w, d) tokens (ctvit.py:306-331), forward only"""
        b, t, h, w, d = tokens.shape
        W = self.weights()
        x = tokens.reshape(-1, d).contiguous().float()
        grid = (b, t, h, w)
        with torch.no_grad():
            tab, rowmax, _ = engine.bias_tables(self, h, w, x.device)
            for L in W.spatial:
                x, _ = engine.layer_forward(x, L, grid, W.heads, False, tab, rowmax, False)
            _, _, x = ops.layernorm_fwd(x, W.s_out, None, want_bf16=False, want_f32=True)
            for L in W.temporal:
                x, _ = engine.layer_forward(x, L, grid, W.heads, True, None, None, False)
            _, _, x = ops.layernorm_fwd(x, W.t_out, None, want_bf16=False, want_f32=True)
        return x.reshape(b, t, h, w, d)

    def encode_pooled(self, video):
        """mean over t of the quantised tokens, flattened to (b, h*w*d) — the tensor CTCLIP projects (ct_clip.py:724,740)"""
        return EncodeFunction.apply(self, video, "pooled", *self.parameters())

    def forward(self, video, mask=None, return_recons=False, return_recons_only=False, return_discr_loss=False,
                apply_grad_penalty=True, return_only_codebook_ids=False, return_encoded_tokens=False):
        assert video.ndim == 5, "CT volumes are (b, 1, f, h, w)"
        if mask is not None:
            raise NotImplementedError("frame masks are not used by CT-CLIP and not implemented")
        b, c, f, *image_dims = video.shape
        assert tuple(image_dims) == self.image_size
        if not (return_only_codebook_ids or return_encoded_tokens):
            raise NotImplementedError("the reconstruction / GAN half of CTViT (ctvit.py:438-546) is out of scope of this "
                                      "implementation; use return_encoded_tokens=True or return_only_codebook_ids=True")
        tokens = EncodeFunction.apply(self, video, "tokens", *self.parameters())
        if return_only_codebook_ids:
            return self.last_indices.reshape(b, -1).long().view(b, *self.last_indices.shape[1:])
        return tokens
