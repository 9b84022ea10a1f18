"""Training step of the reference trainer (CTPA_CLIP/ct_clip/CTCLIPTrainer.py:316-353) on one process per GPU:
forward(loss) -> backward -> gradient all-reduce (NCCL over NVLink) -> clip_grad_norm_(0.5) -> Adam(lr, betas (0.9, 0.99)).

All trainable parameters live in ONE flat fp32 arena (parameters are views into it), with matching flat gradient /
Adam-moment arenas, so clipping and the optimiser are two HBM passes and the all-reduce is a handful of large buckets.
"""
from __future__ import annotations

import os
from pathlib import Path

import torch
import torch.distributed as dist

from . import ops, shadow


# parameters that exist in the reference's state_dict but never receive a gradient on the CLIP path (SURVEY.md a16;
# the reference needs DDP(find_unused_parameters=True) for them, CTCLIPTrainer.py:213): static, so excluded up front
UNUSED_PARAM_MARKERS = ("_latent_extra", "to_pixels", "to_patch_emb_first_frame", "context_norm", "null_kv",
                        "text_transformer.pooler")


def trainable_parameters(module):
    return [(n, p) for n, p in module.named_parameters()
            if p.requires_grad and p.numel() > 0 and not any(m in n for m in UNUSED_PARAM_MARKERS)]


def _arena_order(named):
    """parameter order inside the arena: as named_parameters(), except that the query / key / value projections of a BERT
    self-attention are laid out back to back (weights, then biases), so that the packed [3*hidden, hidden] QKV weight
    gradient of the native text tower is ONE contiguous view of the gradient arena"""
    by_name = dict(named)
    out, done = [], set()
    for n, p in named:
        if n in done:
            continue
        if n.endswith("attention.self.query.weight"):
            base = n[: -len("query.weight")]
            group = [base + f"{m}.{k}" for k in ("weight", "bias") for m in ("query", "key", "value")]
            if all(g in by_name for g in group):
                for g in group:
                    out.append(by_name[g])
                    done.add(g)
                continue
        out.append(p)
        done.add(n)
    return out


class ParamArena:
    def __init__(self, module: torch.nn.Module):
        params = _arena_order(trainable_parameters(module))
        self.params = params
        dev = params[0].device
        sizes = [(p.numel() + 7) // 8 * 8 for p in params]   # 32-byte fp32 slots: the bf16 mirror views stay 16-byte aligned (TMA)
        total = sum(sizes)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.v = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        self.span = {}                    # id(parameter) -> (offset, padded size) inside the arenas
        with torch.no_grad():
            for p, s in zip(params, sizes):
                view = self.flat[off: off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[off: off + p.numel()].view(p.shape)
                self.span[id(p)] = (off, s)
                off += s
        self.norm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.total = total
        # bf16 mirror of the parameters (same offsets): written by the Adam kernel, read by the tensor-core GEMMs as
        # zero-copy views (shadow.bf16_of). `versions` pins the torch-side version of each parameter the mirror matches.
        self.bf16 = None
        self.versions = {}
        if dev.type == "cuda":
            self.bf16 = torch.empty(total, device=dev, dtype=torch.bfloat16)
            self.sync_shadow()
            shadow.register(self)

    def sync_shadow(self):
        """re-derive the whole mirror from the fp32 parameters (construction, or after torch-side writes such as load_state_dict)"""
        if self.bf16 is None:
            return
        ops.cast_bf16(self.flat, out=self.bf16)
        self.versions = {p.data_ptr(): p._version for p in self.params}


class CTClipTrainStep:
    """One optimisation step with the reference's hyper-parameters (CTCLIPTrainer.py:203-205: lr 1.25e-6, wd 0,
    max_grad_norm 0.5; optimizer.py:13-14: betas (0.9, 0.99), eps 1e-8)."""

    def __init__(self, model, lr=1.25e-6, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=0.5, bucket_elems=64 * 1024 * 1024):
        self.model = model
        self.lr, self.betas, self.eps, self.max_grad_norm = lr, betas, eps, max_grad_norm
        self.arena = ParamArena(model)
        # every trainable parameter now owns a zeroed .grad view of the flat arena: let the kernels accumulate into it
        model.direct_grad = True
        model.visual_transformer.direct_grad = True
        self.step_count = 0
        self.bucket = bucket_elems
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self._pending, self._reduced = [], []
        # incremented by the optimiser kernel whenever it refuses an update because the gradient norm is not finite
        self.skipped = torch.zeros(1, device=self.arena.flat.device, dtype=torch.int32)
        if self.distributed:
            self.broadcast_from_rank0()
            vit = model.visual_transformer

            def ema_reduce(bins, esum):
                # bins and embed_sum are two views of ONE buffer (ops.vq_ema): one asynchronous all-reduce for both
                buf = torch.as_strided(bins, (bins.numel() + esum.numel(),), (1,))
                return dist.all_reduce(buf, async_op=True)
            vit.ema_reduce = ema_reduce
            # gradients of a finished image-tower layer start their all-reduce while the earlier layers still back-propagate
            vit.grad_ready = self._on_grads_ready
            # Overlap: the latent-projection and text-tower gradients are final long before the image tower's backward ends
            # (its 294912 -> 512 projection is the FIRST thing the backward computes, the text tower runs next); their
            # all-reduce (1.04 of the 1.13 GB) starts the moment the kernels that wrote them are enqueued.
            self.overlap = os.environ.get("CTCLIP_DP_OVERLAP", "1") != "0"   # 0: ONE gradient all-reduce pass after backward (A/B)
            model.grad_ready = self._on_grads_ready
            # the 294912 -> 512 projection's gradient (604 of the 1130 MB) is rank-B_glob: its bf16 factors are all-gathered
            # (4.7 MB per rank) and multiplied locally instead of all-reducing the product (CTCLIP_FACTOR_GATHER=0: all-reduce)
            model.factor_gather = os.environ.get("CTCLIP_FACTOR_GATHER", "1") != "0"

    def broadcast_from_rank0(self):
        """Replicas must start bit-identical: gradient summation, the factor gather and the EMA all-reduce all assume it. The
        reference gets this from the DDP wrap (parameters broadcast at construction, buffers before every forward,
        CTCLIPTrainer.py:213-217); here the flat parameter arena and every buffer (VQ codebook, cluster_size, ...) are
        broadcast once — the EMA statistics are all-reduced afterwards, so the buffers stay identical — and the derived
        operands are rebuilt. A checksum across ranks confirms it."""
        a = self.arena
        dist.broadcast(a.flat, 0)
        for b in self.model.buffers():
            if b.is_floating_point() or b.dtype in (torch.int64, torch.int32, torch.bool):
                dist.broadcast(b, 0)
        a.sync_shadow()
        self._invalidate_derived()
        chk = torch.stack([a.flat.double().sum(), a.flat.double().abs().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            raise RuntimeError("CTClipTrainStep: parameter replicas differ after the rank-0 broadcast")

    def _invalidate_derived(self):
        """parameters / buffers changed behind the operand caches' back: drop them"""
        self.model.visual_transformer.invalidate_weights()
        self.model._sh_text.key = None
        self.model._sh_vis.key = None
        if getattr(self.model, "_native_text", None) is not None:
            self.model._native_text.invalidate()

    def raise_if_skipped(self):
        """synchronises: raises if the optimiser kernel ever refused an update (non-finite gradient norm) or the latent
        exchange gave up waiting for a peer. Call it wherever the host already syncs (logging, checkpoints)."""
        from . import symm
        n = int(self.skipped.item())
        peer = self.distributed and symm.timed_out()
        if n or peer:
            raise RuntimeError(f"CTClipTrainStep: {n} optimiser update(s) skipped because the gradient norm was not finite"
                               + (" — the latent exchange timed out waiting for a peer rank" if peer else "")
                               + "; parameters and Adam moments were left untouched")

    def _on_grads_ready(self, params, reduced=False):
        """called from inside backward (direct-gradient mode): these parameters' .grad in the arena is final;
        reduced=True: it already is the sum over ranks (factor gather), so the span is only excluded from the all-reduce"""
        if not self.overlap and not reduced:
            return                                 # everything is reduced after backward (reduce_gradients)
        spans = sorted(self.arena.span[id(p)] for p in params if id(p) in self.arena.span)
        merged = []
        for off, n in spans:
            if merged and merged[-1][0] + merged[-1][1] == off:
                merged[-1][1] += n
            else:
                merged.append([off, n])
        for off, n in merged:
            if reduced:
                self._reduced.append((off, off + n))
                continue
            for o in range(off, off + n, self.bucket):
                e = min(off + n, o + self.bucket)
                self._pending.append(dist.all_reduce(self.arena.grad[o:e], async_op=True))
            self._reduced.append((off, off + n))

    def forward_backward(self, text, video):
        if not self.model.training:       # nn.Module.train() walks ~600 modules: only when the mode actually changes
            self.model.train()
        loss = self.model(text, video, return_loss=True)
        loss.backward()
        return loss

    def reduce_gradients(self):
        """sum over ranks: every rank back-propagated its own rows of the global-batch loss"""
        self.model.visual_transformer.finish_ema()    # the (all-reduced) EMA codebook update of this step's forward
        if not self.distributed:
            return
        g = self.arena.grad
        pos = 0
        for a, b in sorted(self._reduced) + [(g.numel(), g.numel())]:   # everything not reduced from inside backward
            for off in range(pos, a, self.bucket):
                self._pending.append(dist.all_reduce(g[off: min(a, off + self.bucket)], async_op=True))
            pos = max(pos, b)
        for w in self._pending:
            w.wait()
        self._pending, self._reduced = [], []

    def optimizer_step(self):
        a = self.arena
        self.step_count += 1
        a.norm_sq.zero_()
        ops.sumsq(a.grad, a.norm_sq)
        ops.adam_step(a.flat, a.grad, a.m, a.v, a.bf16, self.lr, self.betas[0], self.betas[1], self.eps, self.step_count,
                      norm_sq=a.norm_sq, max_norm=self.max_grad_norm, zero_grad=True, skipped=self.skipped)
        # parameters changed in place behind autograd's back: drop the derived operand caches
        self._invalidate_derived()

    def step(self, text, video):
        loss = self.forward_backward(text, video)
        self.reduce_gradients()
        self.optimizer_step()
        return loss

    # ---------------------------------------------------------------- checkpoint I/O (CTCLIPTrainer.py:289-307)
    def save(self, path):
        """Same package layout as CTClipTrainer.save: dict(model=<state_dict>, optim=<optimiser state>), written by rank 0;
        all ranks leave together (barrier), so no rank runs ahead into the next step's latent exchange while rank 0 writes.
        The model part loads with `CTCLIP.load` (strict=False, ct_clip.py:593-597) here and in the reference; a STRICT load
        into a reference CTCLIP built with use_vgg_and_gan=True (pretrained_model.py) additionally expects the `vgg.*` /
        `discr.*` keys of the reconstruction half, which this implementation never creates (out of scope). The optimiser part is
        keyed by parameter NAME (exp_avg / exp_avg_sq / step) because the flat-arena Adam has no torch param-group numbering."""
        self.raise_if_skipped()
        if self.distributed and dist.get_rank() != 0:
            dist.barrier()
            return
        a = self.arena
        names = {id(p): n for n, p in self.model.named_parameters()}
        exp_avg, exp_avg_sq = {}, {}
        for p in a.params:
            off, _ = a.span[id(p)]
            exp_avg[names[id(p)]] = a.m[off: off + p.numel()].view(p.shape).clone()
            exp_avg_sq[names[id(p)]] = a.v[off: off + p.numel()].view(p.shape).clone()
        pkg = dict(model=self.model.state_dict(),
                   optim=dict(format="ctpa_clip_b200.flat_adam", step=self.step_count, exp_avg=exp_avg, exp_avg_sq=exp_avg_sq,
                              lr=self.lr, betas=self.betas, eps=self.eps, max_grad_norm=self.max_grad_norm))
        nt = self.model.ensure_native_text()
        if nt is not None:
            pkg["optim"]["text_dropout_rng"] = nt.rng_state()
        torch.save(pkg, str(path))
        if self.distributed:
            dist.barrier()

    def load(self, path):
        """CTClipTrainer.load: model state_dict (parameters stay views of the arena: load_state_dict copies in place), then
        the Adam moments when the package was written by `save`; a torch.optim state (reference trainer) restores the model
        only. Derived operands (bf16 mirror, cached bf16 / padded weights) are re-synchronised."""
        path = Path(path)
        assert path.exists()
        pkg = torch.load(str(path), map_location=self.arena.flat.device)
        state = pkg["model"] if isinstance(pkg, dict) and "model" in pkg else pkg     # CTCLIP.load takes the bare state_dict
        self.model.load_state_dict(state, strict=False)
        opt = pkg.get("optim") if isinstance(pkg, dict) else None
        if isinstance(opt, dict) and opt.get("format") == "ctpa_clip_b200.flat_adam":
            a = self.arena
            names = {id(p): n for n, p in self.model.named_parameters()}
            with torch.no_grad():
                for p in a.params:
                    off, _ = a.span[id(p)]
                    n = names[id(p)]
                    if n in opt["exp_avg"]:
                        a.m[off: off + p.numel()].view(p.shape).copy_(opt["exp_avg"][n])
                        a.v[off: off + p.numel()].view(p.shape).copy_(opt["exp_avg_sq"][n])
            self.step_count = int(opt["step"])
            if "text_dropout_rng" in opt:
                nt = self.model.ensure_native_text()
                if nt is not None:
                    nt.load_rng_state(opt["text_dropout_rng"])
        self.arena.sync_shadow()
        self._invalidate_derived()
