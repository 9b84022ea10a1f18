"""The injected BERT text tower of CT-CLIP on the sm_100a kernels.

Reference: `CTCLIP.forward` calls `self.text_transformer(text.input_ids, attention_mask=text.attention_mask)[0]`
(ct_clip.py:685-686) on the HF `BertModel` injected by pretrained_model.py:9 (microsoft/BiomedVLP-CXR-BERT-specialized:
BERT-base, post-LayerNorm, absolute position embeddings, erf GELU, LayerNorm eps 1e-12, dropout 0.1 in train mode).

`NativeBert` reads the parameters of that *unchanged* HF module (state_dict keys stay HF's), runs the forward and the
backward through libctclip_sm100.so — tcgen05 GEMMs for every projection and for the per-head Q K^T / P V products
(batched mode), memory-bound kernels for embeddings / softmax / GELU / LayerNorm / dropout — and hands the parameter
gradients back to autograd. There is no fallback: unsupported BERT variants are reported by `supports()` and the caller
keeps the HF module (a library path) for them.
"""
from __future__ import annotations

import math
import os

import torch

from .. import ops
from ..shadow import bf16_of, bf16_rows_of, f32_cat_of


def supports(module) -> bool:
    """True for a transformers BertModel with the architecture CT-CLIP uses (absolute positions, gelu, encoder only)"""
    cfg = getattr(module, "config", None)
    if cfg is None or module.__class__.__name__ != "BertModel":
        return False
    if getattr(cfg, "position_embedding_type", "absolute") != "absolute" or getattr(cfg, "is_decoder", False):
        return False
    if getattr(cfg, "hidden_act", "gelu") != "gelu" or getattr(cfg, "add_cross_attention", False):
        return False
    hd = cfg.hidden_size // cfg.num_attention_heads
    return cfg.hidden_size % 8 == 0 and hd % 8 == 0 and cfg.intermediate_size % 8 == 0 and cfg.hidden_size <= 1024


class _LayerW:
    pass


class NativeBert:
    def __init__(self, module):
        assert supports(module)
        self.m = module
        cfg = module.config
        self.D, self.H, self.I = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        self.hd = self.D // self.H
        self.eps = cfg.layer_norm_eps
        self.p_hidden, self.p_attn = cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob
        self.vocab = cfg.vocab_size
        self._key, self._layers = None, None
        # train-mode dropout masks are a stateless hash of (seed, element): the per-call seed mixes torch's seed at
        # construction (so torch.manual_seed controls it), the data-parallel rank (ranks draw different masks, as DDP replicas
        # do) and the call counter; (seed_base, calls) travel with the trainer checkpoint so a resumed run continues the
        # sequence instead of replaying it from step 0
        self.seed_base = int(torch.initial_seed()) & 0x7FFFFFFF
        self.calls = 0
        # fused attention core (csrc/bert_attn.cu) for head dim 64; CTCLIP_BERT_FUSED_ATTN=0 selects the GEMM + softmax path
        self.fused_attention = self.hd == 64 and os.environ.get("CTCLIP_BERT_FUSED_ATTN", "1") != "0"
        # True (set by CTClipTrainStep): the kernels accumulate parameter gradients straight into the existing p.grad
        # buffers (flat gradient arena, zeroed by the optimiser kernel) and autograd gets None for them
        self.direct_grad = False
        self.grad_ready = None      # optional callable(list of parameters), called at the end of a direct-mode backward
        # parameters in a fixed order (the autograd.Function takes them as inputs and returns their gradients)
        emb = module.embeddings
        self.names, self.params = [], []
        for n, p in module.named_parameters():
            if n.startswith("pooler."):
                continue  # never reached: CT-CLIP reads last_hidden_state[:, 0] (ct_clip.py:762), not pooler_output
            self.names.append(n)
            self.params.append(p)
        self.index = {n: i for i, n in enumerate(self.names)}
        assert emb.word_embeddings.weight is self.params[self.index["embeddings.word_embeddings.weight"]]

    # bf16 operand copies (derived caches, rebuilt when a parameter changes)
    def layers(self):
        key = tuple((p.data_ptr(), p._version) for p in self.params[:8]) + (self.params[-1]._version,)
        if self._layers is not None and key == self._key:
            return self._layers
        out = []
        for layer in self.m.encoder.layer:
            a, w = layer.attention, _LayerW()
            s = a.self
            w.wqkv = bf16_rows_of((s.query.weight, s.key.weight, s.value.weight))
            w.bqkv = f32_cat_of((s.query.bias, s.key.bias, s.value.bias))
            w.wo = bf16_of(a.output.dense.weight)
            w.bo = a.output.dense.bias.detach()
            w.g1, w.b1n = a.output.LayerNorm.weight.detach(), a.output.LayerNorm.bias.detach()
            w.w1 = bf16_of(layer.intermediate.dense.weight)
            w.b1 = layer.intermediate.dense.bias.detach()
            w.w2 = bf16_of(layer.output.dense.weight)
            w.b2 = layer.output.dense.bias.detach()
            w.g2, w.b2n = layer.output.LayerNorm.weight.detach(), layer.output.LayerNorm.bias.detach()
            out.append(w)
        self._key, self._layers = key, out
        return out

    def invalidate(self):
        self._key = None

    def _call_seed(self) -> int:
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        return (self.seed_base * 1000003 + rank * 15485863 + self.calls * 7919 + 17) & 0x7FFFFFFF

    def rng_state(self) -> dict:
        return {"seed_base": self.seed_base, "calls": self.calls}

    def load_rng_state(self, st: dict):
        self.seed_base, self.calls = int(st["seed_base"]), int(st["calls"])

    def _packed_qkv_grad(self, pfx):
        """([3D, D] weight-gradient view, [3D] bias-gradient view) when the q / k / v .grad buffers of this layer sit back
        to back in memory (trainer.ParamArena lays them out that way), else None"""
        D = self.D
        ws = [self.params[self.index[pfx + f"attention.self.{m}.weight"]].grad for m in ("query", "key", "value")]
        bs = [self.params[self.index[pfx + f"attention.self.{m}.bias"]].grad for m in ("query", "key", "value")]
        if any(g is None or g.dtype != torch.float32 or not g.is_contiguous() for g in ws + bs):
            return None
        if not (ws[1].data_ptr() == ws[0].data_ptr() + 4 * D * D and ws[2].data_ptr() == ws[1].data_ptr() + 4 * D * D and
                bs[1].data_ptr() == bs[0].data_ptr() + 4 * D and bs[2].data_ptr() == bs[1].data_ptr() + 4 * D):
            return None
        if ws[0].untyped_storage().data_ptr() != ws[2].untyped_storage().data_ptr():
            return None
        return ws[0].as_strided((3 * D, D), (D, 1)), bs[0].as_strided((3 * D,), (1,))

    # ------------------------------------------------------------------------------------------ attention products
    def _scores(self, qkv, B, L):
        """S[b,h] = Q_bh K_bh^T  (fp32 [B,H,L,L])"""
        D, H, hd = self.D, self.H, self.hd
        S = torch.empty((B * H, L, L), device=qkv.device, dtype=torch.float32)
        base = qkv.data_ptr()
        d = ops.gemm_batched(base, 3 * D, False, hd, L * 3 * D, base + 2 * D, 3 * D, False, hd, L * 3 * D,
                             S, L, L * L, H * L * L, L, L, hd, H, B)
        ops.run_gemm_desc(d, True, f"bert_qk:{B}x{H}x{L}x{L}x{hd}", 2.0 * B * H * L * L * hd)
        return S

    def _pv(self, P, qkv, B, L):
        """ctx[b, :, h] = P_bh V_bh  (bf16 [B*L, D])"""
        D, H, hd = self.D, self.H, self.hd
        ctx = torch.empty((B * L, D), device=qkv.device, dtype=torch.bfloat16)
        d = ops.gemm_batched(P.data_ptr(), L, False, L * L, H * L * L, qkv.data_ptr() + 2 * (2 * D), 3 * D, True, hd,
                             L * 3 * D, ctx, D, hd, L * D, L, hd, L, H, B)
        ops.run_gemm_desc(d, False, f"bert_pv:{B}x{H}x{L}x{hd}x{L}", 2.0 * B * H * L * L * hd)
        return ctx

    def _attention_backward_unfused(self, qkv, P, Pd, dcv, B, L, scale, pa, sd, dev):
        """BertSelfAttention backward as batched tcgen05 GEMMs + the softmax-backward kernel (any head dim)"""
        D, H, hd = self.D, self.H, self.hd
        Pop = Pd if Pd is not None else P
        base = qkv.data_ptr()
        dqkv = torch.empty_like(qkv)
        dbase = dqkv.data_ptr()
        fl = 2.0 * B * H * L * L * hd
        # dP = dctx V^T   (fp32 [B,H,L,L])
        dP = torch.empty((B * H, L, L), device=dev, dtype=torch.float32)
        d = ops.gemm_batched(dcv.data_ptr(), D, False, hd, L * D, base + 2 * (2 * D), 3 * D, False, hd, L * 3 * D,
                             dP, L, L * L, H * L * L, L, L, hd, H, B)
        ops.run_gemm_desc(d, True, f"bert_dp:{B}x{H}x{L}x{L}x{hd}", fl)
        # dV = P^T dctx  -> v slice of dqkv
        d = ops.gemm_batched(Pop.data_ptr(), L, True, L * L, H * L * L, dcv.data_ptr(), D, True, hd, L * D,
                             dbase + 2 * (2 * D), 3 * D, hd, L * 3 * D, L, hd, L, H, B)
        ops.run_gemm_desc(d, False, f"bert_dv:{B}x{H}x{L}x{hd}x{L}", fl)
        dS = ops.bert_softmax_bwd(P, dP, B, H, L, scale, pa, sd + 1)
        del dP
        # dQ = dS K -> q slice ; dK = dS^T Q -> k slice
        d = ops.gemm_batched(dS.data_ptr(), L, False, L * L, H * L * L, base + 2 * D, 3 * D, True, hd, L * 3 * D,
                             dbase, 3 * D, hd, L * 3 * D, L, hd, L, H, B)
        ops.run_gemm_desc(d, False, f"bert_dq:{B}x{H}x{L}x{hd}x{L}", fl)
        d = ops.gemm_batched(dS.data_ptr(), L, True, L * L, H * L * L, base, 3 * D, True, hd, L * 3 * D,
                             dbase + 2 * D, 3 * D, hd, L * 3 * D, L, hd, L, H, B)
        ops.run_gemm_desc(d, False, f"bert_dk:{B}x{H}x{L}x{hd}x{L}", fl)
        return dqkv

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, ids, mask, training: bool):
        B, L = ids.shape
        D, H, hd, I = self.D, self.H, self.hd, self.I
        T = B * L
        emb = self.m.embeddings
        ph = self.p_hidden if training else 0.0
        pa = self.p_attn if training else 0.0
        self.calls += 1
        seed0 = self._call_seed()
        ids = ids.contiguous()
        mask = mask.contiguous().to(torch.long)
        ctx = {"B": B, "L": L, "ph": ph, "pa": pa, "seed0": seed0, "ids": ids, "mask": mask, "layers": []}
        pre0 = ops.bert_embed_fwd(ids.reshape(-1), L, emb.word_embeddings.weight.detach(),
                                  emb.position_embeddings.weight.detach(),
                                  emb.token_type_embeddings.weight.detach()[0].contiguous())
        xb, _, x = ops.layernorm_fwd(pre0, emb.LayerNorm.weight.detach(), emb.LayerNorm.bias.detach(), eps=self.eps,
                                     want_bf16=(ph == 0), want_f32=True)
        if ph > 0:
            x = ops.dropout_add(x, None, ph, seed0)
            xb = ops.cast_bf16(x)
        ctx["pre0"] = pre0
        scale = 1.0 / math.sqrt(hd)
        for li, w in enumerate(self.layers()):
            c = {}
            sd = seed0 + 101 * (li + 1)
            qkv = ops.gemm(xb, w.wqkv, bias=w.bqkv, tag=f"bert_qkv:{T}x{3 * D}x{D}")
            if self.fused_attention:    # scores / probabilities stay in registers (csrc/bert_attn.cu)
                cv, P = ops.bert_attn_fwd(qkv, mask, B, H, L, D, pa, sd + 1)     # "P" slot holds lse
                Pd = None
            else:
                S = self._scores(qkv, B, L)
                P, Pd = ops.bert_softmax_fwd(S, mask, B, H, L, scale, pa, sd + 1)
                del S
                cv = self._pv(Pd if Pd is not None else P, qkv, B, L)
            if ph > 0:
                ao = ops.gemm(cv, w.wo, out_dtype=torch.float32, bias=w.bo, tag=f"bert_out:{T}x{D}x{D}")
                pre1 = ops.dropout_add(ao, x, ph, sd + 2)
            else:
                pre1 = ops.gemm(cv, w.wo, out_dtype=torch.float32, bias=w.bo, resid=x, tag=f"bert_out:{T}x{D}x{D}")
            x1b, _, x1 = ops.layernorm_fwd(pre1, w.g1, w.b1n, eps=self.eps, want_bf16=True, want_f32=True)
            h = ops.gemm(x1b, w.w1, bias=w.b1, tag=f"bert_ff1:{T}x{I}x{D}")
            a = ops.gelu_fwd(h)
            if ph > 0:
                y = ops.gemm(a, w.w2, out_dtype=torch.float32, bias=w.b2, tag=f"bert_ff2:{T}x{D}x{I}")
                pre2 = ops.dropout_add(y, x1, ph, sd + 3)
            else:
                pre2 = ops.gemm(a, w.w2, out_dtype=torch.float32, bias=w.b2, resid=x1, tag=f"bert_ff2:{T}x{D}x{I}")
            x2b, _, x2 = ops.layernorm_fwd(pre2, w.g2, w.b2n, eps=self.eps, want_bf16=True, want_f32=True)
            c["xb"], c["qkv"], c["P"], c["Pd"], c["cv"], c["pre1"], c["x1b"], c["h"], c["a"], c["pre2"], c["sd"] = \
                xb, qkv, P, Pd, cv, pre1, x1b, h, a, pre2, sd
            ctx["layers"].append(c)
            x, xb = x2, x2b
        return x.view(B, L, D), ctx

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, ctx, g_out):
        """g_out: dL/d(last_hidden_state) fp32 [B, L, D]. Returns the list of parameter gradients in self.params order."""
        B, L, ph, pa = ctx["B"], ctx["L"], ctx["ph"], ctx["pa"]
        D, H, hd, I = self.D, self.H, self.hd, self.I
        T = B * L
        dev = g_out.device
        grads = [None] * len(self.params)
        direct = self.direct_grad

        def target(name):
            """fp32 accumulator for the gradient of parameter `name`: p.grad itself in direct mode, else a fresh zero
            tensor that is handed to autograd"""
            i = self.index[name]
            pg = self.params[i].grad
            if direct and pg is not None and pg.is_contiguous() and pg.dtype == torch.float32:
                return pg
            t = torch.zeros(self.params[i].shape, device=dev, dtype=torch.float32)
            grads[i] = t
            return t

        def wgrad(name, dy_bf, x_bf):
            dw = target(name)
            ops.gemm(dy_bf, x_bf, a_t=True, b_t=True, out=dw, accumulate=True, splits=0,
                     tag=f"bert_wgrad:{dw.shape[0]}x{dw.shape[1]}x{T}")

        scale = 1.0 / math.sqrt(hd)
        g = g_out.reshape(T, D).contiguous().float()
        layers = self.layers()
        for li in reversed(range(len(layers))):
            w, c = layers[li], ctx["layers"][li]
            sd = c["sd"]
            pfx = f"encoder.layer.{li}."
            # ---- BertOutput: LayerNorm(dropout(dense(a)) + x1)
            dpre2, dpre2_b = ops.layernorm_bwd(g, c["pre2"], w.g2, eps=self.eps, dgamma=target(pfx + "output.LayerNorm.weight"),
                                               dbeta=target(pfx + "output.LayerNorm.bias"), want_bf16=(ph == 0))
            if ph > 0:
                dy = ops.dropout_add(dpre2, None, ph, sd + 3)
                dy_b = ops.cast_bf16(dy)
            else:
                dy, dy_b = dpre2, dpre2_b
            ops.colsum(dy, target(pfx + "output.dense.bias"))
            wgrad(pfx + "output.dense.weight", dy_b, c["a"])
            da = ops.gemm(dy_b, w.w2, b_t=True, tag=f"bert_dgrad:{T}x{I}x{D}")
            # ---- BertIntermediate: gelu(dense(x1))
            dh = ops.gelu_bwd(c["h"], da)
            ops.colsum_bf16(dh, target(pfx + "intermediate.dense.bias"))
            wgrad(pfx + "intermediate.dense.weight", dh, c["x1b"])
            # x1 feeds the FFN and, as the residual, pre2: its gradient is the FFN data gradient + dpre2 (GEMM epilogue)
            dx1 = ops.gemm(dh, w.w1, b_t=True, out_dtype=torch.float32, resid=dpre2, tag=f"bert_dgrad:{T}x{D}x{I}")
            del dh, da
            # ---- BertSelfOutput: LayerNorm(dropout(dense(ctx)) + x)
            dpre1, dpre1_b = ops.layernorm_bwd(dx1, c["pre1"], w.g1, eps=self.eps,
                                               dgamma=target(pfx + "attention.output.LayerNorm.weight"),
                                               dbeta=target(pfx + "attention.output.LayerNorm.bias"), want_bf16=(ph == 0))
            if ph > 0:
                dao = ops.dropout_add(dpre1, None, ph, sd + 2)
                dao_b = ops.cast_bf16(dao)
            else:
                dao, dao_b = dpre1, dpre1_b
            ops.colsum(dao, target(pfx + "attention.output.dense.bias"))
            wgrad(pfx + "attention.output.dense.weight", dao_b, c["cv"])
            dcv = ops.gemm(dao_b, w.wo, b_t=True, tag=f"bert_dgrad:{T}x{D}x{D}")
            # ---- BertSelfAttention
            qkv, P, Pd = c["qkv"], c["P"], c["Pd"]
            if self.fused_attention:
                dqkv = ops.bert_attn_bwd(qkv, ctx["mask"], c["cv"], P, dcv, B, H, L, D, pa, sd + 1)
            else:
                dqkv = self._attention_backward_unfused(qkv, P, Pd, dcv, B, L, scale, pa, sd, dev)
            packed = self._packed_qkv_grad(pfx) if direct else None
            if packed is not None:   # q / k / v gradients are adjacent in the arena: one wgrad GEMM, one column sum
                dw3, db3 = packed
                ops.colsum_bf16(dqkv, db3)
                ops.gemm(dqkv, c["xb"], a_t=True, b_t=True, out=dw3, accumulate=True, splits=0,
                         tag=f"bert_wgrad:{3 * D}x{D}x{T}")
            else:
                for j, nm in enumerate(("query", "key", "value")):   # the three projections are separate parameters
                    dslice = dqkv[:, j * D:(j + 1) * D]                # column slice of the packed gradient (row pitch 3D)
                    ops.colsum_bf16(dslice, target(pfx + f"attention.self.{nm}.bias"), dim=D, ld=3 * D)
                    wgrad(pfx + f"attention.self.{nm}.weight", dslice, c["xb"])
            g = ops.gemm(dqkv, w.wqkv, b_t=True, out_dtype=torch.float32, resid=dpre1, tag=f"bert_dgrad:{T}x{D}x{3 * D}")
            ctx["layers"][li] = None
        # ---- embeddings: LayerNorm(word + pos + type) (+ dropout)
        emb = self.m.embeddings
        if ph > 0:
            g = ops.dropout_add(g, None, ph, ctx["seed0"])
        dpre0, _ = ops.layernorm_bwd(g, ctx["pre0"], emb.LayerNorm.weight.detach(), eps=self.eps,
                                     dgamma=target("embeddings.LayerNorm.weight"), dbeta=target("embeddings.LayerNorm.bias"))
        ops.bert_embed_bwd(ctx["ids"].reshape(-1), L, dpre0, target("embeddings.word_embeddings.weight"),
                           target("embeddings.position_embeddings.weight"), emb.word_embeddings.padding_idx)
        ops.colsum(dpre0, target("embeddings.token_type_embeddings.weight")[0])
        if direct and self.grad_ready is not None and all(g is None for g in grads):
            self.grad_ready(self.params)   # every gradient went straight into p.grad and is final
        return grads


class BertFunction(torch.autograd.Function):
    """last_hidden_state = BertModel(input_ids, attention_mask)[0] with a hand-written backward"""

    @staticmethod
    def forward(ctx, engine, ids, mask, training, *params):
        out, saved = engine.forward(ids, mask, training)
        ctx.engine, ctx.saved = engine, saved
        return out

    @staticmethod
    def backward(ctx, g):
        grads = ctx.engine.backward(ctx.saved, g)
        ctx.saved = None
        return (None, None, None, None, *grads)


def encode(engine: NativeBert, ids, mask, training: bool):
    return BertFunction.apply(engine, ids, mask, training, *engine.params)
