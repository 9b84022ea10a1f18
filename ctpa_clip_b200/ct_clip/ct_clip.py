"""CTCLIP — same constructor, forward signature, attributes and state_dict keys as the live branch of
CTPA_CLIP/ct_clip/ct_clip.py:407-901 (image_encoder / text_encoder injected; extra_latent_projection, MLM, visual SSL,
FILIP/DCL and multiview branches are disabled by pretrained_model.py:37-40 and raise here if requested).

The image tower, latent projections, l2norm and the symmetric InfoNCE run in libctclip_sm100.so. With
torch.distributed initialised the loss is the GLOBAL-batch InfoNCE: one kernel pushes every rank's l2-normalised latents
into its peers' symmetric buffers over NVLink and forms the logits (csrc/symm.cu; NCCL all-gather with
CTCLIP_LATENT_EXCHANGE=nccl), and every rank differentiates its own rows; the reference's local-batch
loss is the world_size=1 case.
"""
from __future__ import annotations

import copy
from pathlib import Path

import torch
import torch.distributed as dist
from torch import nn

from .. import ops, symm
from .ctvit import CTViT


def exists(val):
    return val is not None


from ..shadow import bf16_of


class _Shadow:
    """bf16 operand copy of an fp32 weight, refreshed when the parameter changes"""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, w):
        key = (w.data_ptr(), w._version)
        if key != self.key:
            self.val = bf16_of(w)          # a view of the optimiser's bf16 mirror when there is one, else a cast
            self.key = key
        return self.val


class LinearFunction(torch.autograd.Function):
    """y = x W^T (bias-free nn.Linear, ct_clip.py:549,564) on the tcgen05 GEMM; fp32 accumulate, fp32 out"""

    @staticmethod
    def forward(ctx, x, weight, w_bf16, x_bf16, direct=False, ready=None, factor_gather=False):
        xb = x_bf16 if x_bf16 is not None else ops.cast_bf16(x.contiguous().float())
        M = xb.shape[0]
        y = torch.zeros((M, weight.shape[0]), device=x.device, dtype=torch.float32)
        ops.gemm(xb, w_bf16, out=y, accumulate=True, splits=0)  # skinny: split-K over the (long) reduction
        ctx.save_for_backward(xb, w_bf16)
        ctx.need = (x.requires_grad, weight.requires_grad)
        ctx.weight = weight if direct else None
        ctx.ready = ready
        ctx.factor_gather = factor_gather
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        dyb = ops.cast_bf16(dy.contiguous().float())
        dx = dw = None
        if ctx.need[0]:
            dx = ops.gemm(dyb, wb, b_t=True, out_dtype=torch.float32)               # [M, K]
        if ctx.need[1]:
            w = ctx.weight
            if w is not None and w.grad is not None and w.grad.is_contiguous():
                # direct mode: accumulate into the parameter's gradient buffer (no 600 MB temporary + add for the
                # 294912 -> 512 projection); autograd gets None
                # (in-place "+=" through the residual epilogue: plain 128-bit loads / stores instead of 151 M atomics)
                if ctx.factor_gather and dist.is_initialized() and dist.get_world_size() > 1:
                    # dW = dy^T x is rank-B: all-gather the two FACTORS (B_loc x N and B_loc x K bf16, 4.7 MB per rank for
                    # the 294912 -> 512 projection) and form the global-batch gradient locally, instead of all-reducing
                    # the N x K fp32 product (604 MB) — SURVEY 7.2-6 / 8(e). Every rank ends with the identical sum.
                    ws = dist.get_world_size()
                    bl, n, kk = dyb.shape[0], dyb.shape[1], xb.shape[1]
                    packed = torch.empty((bl, n + kk), device=dyb.device, dtype=dyb.dtype)   # [dL | E] rows: ONE all-gather
                    packed[:, :n] = dyb
                    packed[:, n:] = xb
                    gathered = torch.empty((ws * bl, n + kk), device=dyb.device, dtype=dyb.dtype)
                    dist.all_gather_into_tensor(gathered, packed)
                    dy_all, x_all = gathered[:, :n], gathered[:, n:]        # strided views: the GEMM takes a row pitch
                    ops.gemm(dy_all, x_all, a_t=True, b_t=True, out=w.grad, resid=w.grad)
                    if ctx.ready is not None:
                        ctx.ready([w], reduced=True)     # already the sum over ranks: no all-reduce for this span
                else:
                    ops.gemm(dyb, xb, a_t=True, b_t=True, out=w.grad, resid=w.grad)
                    if ctx.ready is not None:
                        ctx.ready([w])   # this gradient is final: the trainer may start its all-reduce now
            else:
                dw = ops.gemm(dyb, xb, a_t=True, b_t=True, out_dtype=torch.float32)  # [N, K], reduction over the batch
        return dx, dw, None, None, None, None, None


class ClipLossFunction(torch.autograd.Function):
    """l2norm -> (all-gather) -> exp(tau) T I^T -> symmetric InfoNCE (ct_clip.py:771,796,845-878), with its gradient
    w.r.t. this rank's un-normalised latents and the temperature."""

    @staticmethod
    def forward(ctx, text_raw, image_raw, temperature):
        t_hat, _, t_inv = ops.l2norm_rows(text_raw.contiguous().float(), want_f32=True)
        i_hat, _, i_inv = ops.l2norm_rows(image_raw.contiguous().float(), want_f32=True)
        b = t_hat.shape[0]
        tau = temperature.detach().reshape(1).float().contiguous()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            ws, rank = dist.get_world_size(), dist.get_rank()
            ex = symm.get_exchange(b, t_hat.shape[1])
            if ex is not None:
                # ONE kernel pushes the latents into every peer's buffer over NVLink and forms the global-batch logits
                loss, dT, dI, dtau = ops.clip_loss_allgather(t_hat, i_hat, tau, rank, ws, ex.table, ex.next_step(), ex.status)
            else:
                local = torch.stack((t_hat, i_hat))                                 # [2, b, d]
                gathered = torch.empty((ws, *local.shape), device=local.device, dtype=local.dtype)
                dist.all_gather_into_tensor(gathered, local)                        # 2*b*d fp32 per rank over NVLink
                T = gathered[:, 0].reshape(ws * b, -1).contiguous()
                I = gathered[:, 1].reshape(ws * b, -1).contiguous()
                loss, dT, dI, dtau = ops.clip_loss(T, I, tau, rank * b, b, want_grad=True)
        else:
            loss, dT, dI, dtau = ops.clip_loss(t_hat, i_hat, tau, 0, b, want_grad=True)
        ctx.save_for_backward(t_hat, t_inv, i_hat, i_inv, dT, dI, dtau)
        return loss

    @staticmethod
    def backward(ctx, g):
        t_hat, t_inv, i_hat, i_inv, dT, dI, dtau = ctx.saved_tensors
        dt_raw = ops.l2norm_bwd(t_hat, t_inv, dT)
        di_raw = ops.l2norm_bwd(i_hat, i_inv, dI)
        return dt_raw * g, di_raw * g, (dtau * g).reshape(())


class L2NormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, _, inv = ops.l2norm_rows(x.contiguous().float(), want_f32=True)
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, g):
        y, inv = ctx.saved_tensors
        return ops.l2norm_bwd(y, inv, g.contiguous().float())


class CTCLIP(nn.Module):
    def __init__(self, *, image_encoder=None, text_encoder=None, dim_text=512, dim_image=512, dim_latent=512,
                 num_text_tokens=28897, text_enc_depth=6, text_seq_len=256, text_heads=8, text_dim_head=64,
                 text_has_cls_token=False, text_pad_id=0, text_rotary_pos_emb=False, text_causal_mask=False,
                 text_eos_id=None, text_encode_without_mask=False, visual_enc_depth=6, visual_heads=8,
                 visual_dim_head=64, visual_image_size=256, visual_patch_size=32, visual_patch_dropout=0.5,
                 visual_has_cls_token=False, channels=3, use_all_token_embeds=False, downsample_image_embeds=False,
                 decoupled_contrastive_learning=False, extra_latent_projection=False, use_mlm=False,
                 text_ssl_loss_weight=0.05, use_visual_ssl=False, visual_ssl=None, visual_ssl_type='simsiam',
                 visual_ssl_hidden_layer=-1, simclr_temperature=0.1, image_ssl_loss_weight=0.05,
                 multiview_loss_weight=0.1, checkpoint_during_training=False, tokenizer=None, **kwargs):
        super().__init__()
        if not isinstance(image_encoder, CTViT):
            raise NotImplementedError("image_encoder must be a ctpa_clip_b200 CTViT (the generic VisionTransformer of "
                                      "ct_clip.py:498 is never built by CT-CLIP)")
        if text_encoder is None:
            raise NotImplementedError("text_encoder must be injected (pretrained_model.py:9 injects a HF BertModel)")
        for flag, name in ((use_all_token_embeds, "use_all_token_embeds"), (downsample_image_embeds, "downsample_image_embeds"),
                           (decoupled_contrastive_learning, "decoupled_contrastive_learning"),
                           (extra_latent_projection, "extra_latent_projection"), (use_mlm, "use_mlm"),
                           (use_visual_ssl or exists(visual_ssl), "visual_ssl"), (text_causal_mask, "text_causal_mask")):
            if flag:
                raise NotImplementedError(f"{name}=True is disabled in CT-CLIP (pretrained_model.py:37-40) and not implemented")
        self.dtype = torch.float32
        self.dim_text, self.dim_image, self.dim_latent = dim_text, dim_image, dim_latent
        self.image_channels = channels
        self.image_size = visual_image_size
        self.text_pad_id = text_pad_id
        self.text_has_cls_token = text_has_cls_token
        self.text_seq_len = text_seq_len
        self.text_encode_without_mask = text_encode_without_mask
        self.text_causal_mask = text_causal_mask
        self.text_eos_id = text_eos_id
        self.text_transformer = text_encoder
        self.visual_has_cls_token = visual_has_cls_token
        self.visual_transformer = image_encoder
        self.use_mlm = False
        self.text_ssl_loss_weight = 0
        self.use_visual_ssl = False
        self.image_ssl_loss_weight = 0
        self.to_text_latent = nn.Linear(dim_text, dim_latent, bias=False)
        self.to_visual_latent = nn.Linear(dim_image, dim_latent, bias=False)
        self.temperature = nn.Parameter(torch.tensor(1.))
        self.use_all_token_embeds = False
        self.decoupled_contrastive_learning = False
        self.extra_latent_projection = False
        # dead deep copies kept for state_dict compatibility (ct_clip.py:579-581)
        self.to_text_latent_extra = copy.deepcopy(self.to_text_latent)
        self.to_visual_latent_extra = copy.deepcopy(self.to_visual_latent)
        self.multiview_loss_weight = multiview_loss_weight
        self.tokenizer = tokenizer  # the reference downloads a BertTokenizer here (ct_clip.py:585); inject one instead
        self.text_autocast = True   # non-BERT text encoders: run the injected torch module under bf16 autocast
        self.native_text = True     # HF BertModel (what CT-CLIP injects): forward/backward on libctclip_sm100.so
        self._native_text = None
        # set by CTClipTrainStep: kernels accumulate parameter gradients straight into the existing p.grad buffers
        self.direct_grad = False
        self.grad_ready = None      # optional callable(list of parameters): their .grad is final (direct mode only)
        self.factor_gather = False  # data-parallel: all-gather the factors of the rank-B to_visual_latent gradient (trainer sets it)
        self._sh_text, self._sh_vis = _Shadow(), _Shadow()

    def load(self, path):
        path = Path(path)
        assert path.exists()
        self.load_state_dict(torch.load(str(path)), strict=False)

    # ---------------------------------------------------------------- towers
    def ensure_native_text(self):
        """the NativeBert engine over the injected HF BertModel (None for any other text encoder)"""
        if self._native_text is None and self.native_text:
            from ..text import NativeBert, supports
            if supports(self.text_transformer):
                self._native_text = NativeBert(self.text_transformer)
        return self._native_text

    def encode_text(self, text):
        dev = self.to_text_latent.weight.device
        if self.native_text and dev.type == "cuda":
            from ..text.bert import encode
            if self.ensure_native_text() is not None:   # BERT on the sm_100a kernels (ct_clip.py:685-686)
                self._native_text.direct_grad = self.direct_grad
                self._native_text.grad_ready = self.grad_ready
                training = self.training and self.text_transformer.training and torch.is_grad_enabled()
                return encode(self._native_text, text.input_ids, text.attention_mask, training)
        # any other injected text encoder runs as the torch module it is (library kernels)
        if self.text_autocast and dev.type == "cuda":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                enc_text = self.text_transformer(text.input_ids, attention_mask=text.attention_mask)[0]
        else:
            enc_text = self.text_transformer(text.input_ids, attention_mask=text.attention_mask)[0]
        return enc_text

    def text_latents_raw(self, enc_text):
        cls = enc_text[:, 0, :].float().contiguous()                                # ct_clip.py:762
        return LinearFunction.apply(cls, self.to_text_latent.weight, self._sh_text.get(self.to_text_latent.weight), None,
                                    self.direct_grad, self.grad_ready)

    def image_latents_raw(self, image):
        vit = self.visual_transformer
        pooled = vit.encode_pooled(image)                                           # ct_clip.py:715,724,740
        return LinearFunction.apply(pooled, self.to_visual_latent.weight, self._sh_vis.get(self.to_visual_latent.weight),
                                    vit.last_pooled_bf16, self.direct_grad, self.grad_ready, self.factor_gather)

    # ---------------------------------------------------------------- forward (ct_clip.py:614-901)
    def forward(self, text, image, device=None, return_loss=False, return_encodings=False, return_latents=False,
                freeze_image_encoder=False, freeze_text_encoder=False, text_to_image=True, aug_text=None, aug_image=None):
        if exists(aug_text) or exists(aug_image):
            raise NotImplementedError("multiview augmentation is not used by CT-CLIP and not implemented")
        if return_encodings:
            vit = self.visual_transformer
            return self.encode_text(text), vit.encode_pooled(image)
        # the image tower (few, long kernels) is enqueued first so that the host runs ahead of the GPU while it enqueues
        # the many short kernels of the text tower; the two towers are independent (ct_clip.py:685 / :715)
        image_raw = self.image_latents_raw(image)
        enc_text = self.encode_text(text)
        text_raw = self.text_latents_raw(enc_text)
        if return_loss:
            return ClipLossFunction.apply(text_raw, image_raw, self.temperature)
        text_latents, image_latents = L2NormFunction.apply(text_raw), L2NormFunction.apply(image_raw)
        if return_latents:
            vit = self.visual_transformer
            b, t, h, w = vit.last_indices.shape
            enc_image_send = ops.vq_gather(vit.vq.codebook, vit.last_indices.reshape(-1)).view(b, t, h, w, vit.dim)
            return text_latents, image_latents, enc_image_send
        temp = self.temperature.exp()
        return (text_latents * image_latents).sum(dim=-1) * temp                    # einsum('b d, b d -> b'), ct_clip.py:805-807

    # ---------------------------------------------------------------- batched zero-shot scoring (ctclip_inference.py:286-336)
    @torch.no_grad()
    def prompt_latents(self, prompt_pairs):
        """l2-normalised text latents (2P, d) of tokenised prompt rows [present_0, absent_0, present_1, ...]"""
        return L2NormFunction.apply(self.text_latents_raw(self.encode_text(prompt_pairs)))

    @torch.no_grad()
    def zero_shot_scores(self, prompt_pairs, images, prompt_latents=None):
        """prompt_pairs: tokenised (P*2, L) rows ordered [present_0, absent_0, present_1, ...]; images (V,1,f,h,w).
        Encodes every volume ONCE (the reference re-encodes it, and the prompts, per pathology) and returns the softmax-pair
        prob[present] (V, P). `prompt_latents` (from `prompt_latents()`) skips the text tower for a fixed prompt set."""
        t_lat = prompt_latents if prompt_latents is not None else self.prompt_latents(prompt_pairs)
        i_lat = L2NormFunction.apply(self.image_latents_raw(images))
        tau = self.temperature.detach().reshape(1).float().contiguous()
        return ops.zero_shot_scores(i_lat, t_lat, tau)                              # (V, P) softmax-pair prob[present]
