"""Batched zero-shot evaluator: the inner loop of `CTClipInference.train_step` (CTPA_CLIP/ct_clip/ctclip_inference.py:286-336)
and of the trainer's periodic evaluation (CTCLIPTrainer.py:356-454) on the CUDA engine.

The reference scores one volume at a time and, for each of the 18 pathologies, tokenises the two prompts, runs BERT on
them AND re-runs the image encoder on the same volume (18 image encodes + 36 text encodes per volume). Here the 36 prompt
latents are computed once per evaluation, each volume is encoded once, and `ctclip_zero_shot_scores` forms all
(volume, pathology) softmax pairs. Multi-GPU = replicas: volumes are dealt round-robin to the ranks, the
prompt latents are replicated, the (N, P) score matrix is gathered. The output arrays are what the reference hands to
`evaluate_internal` (evaluate.py:160-207): `predicted_all` (N, P) float and `real_all` (N, P) one-hot labels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

# ctclip_inference.py:297-302 (domain vocabulary: the 18 findings scored by the reference)
PATHOLOGIES = ['Medical material', 'Arterial wall calcification', 'Cardiomegaly', 'Pericardial effusion',
               'Coronary artery wall calcification', 'Hiatal hernia', 'Lymphadenopathy', 'Emphysema', 'Atelectasis',
               'Lung nodule', 'Lung opacity', 'Pulmonary Embolism', 'Pleural effusion', 'Mosaic attenuation pattern',
               'Peribronchial thickening', 'Consolidation', 'Bronchiectasis', 'Interlobular septal thickening']


def prompt_texts(pathologies=PATHOLOGIES):
    """[present_0, absent_0, present_1, ...] exactly as ctclip_inference.py:318 words them"""
    out = []
    for p in pathologies:
        out += [f"{p} is present.", f"{p} is not present."]
    return out


def tokenize_prompts(tokenizer, pathologies=PATHOLOGIES, device="cuda"):
    """ctclip_inference.py:319-320: padding to max_length 512, truncation"""
    return tokenizer(prompt_texts(pathologies), return_tensors="pt", padding="max_length", truncation=True,
                     max_length=512).to(device)


def shard_indices(n: int, rank: int, world: int):
    """round-robin deal of n items: rank r scores items r, r+world, ..."""
    return list(range(rank, n, world))


def gather_rows(local: torch.Tensor, n: int, rank: int, world: int, group=None) -> torch.Tensor:
    """inverse of `shard_indices`: every rank contributes its (len(shard), P) rows, every rank gets the (n, P) matrix in
    dataset order. Shards are padded to a common length so that one all_gather moves everything."""
    if world == 1:
        return local
    per = (n + world - 1) // world
    P = local.shape[1]
    padded = local.new_zeros((per, P))
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = local.new_empty((n, P))
    for r in range(world):
        idx = shard_indices(n, r, world)
        out[idx] = parts[r][: len(idx)]
    return out


class ZeroShotEvaluator:
    def __init__(self, model, prompt_tokens, batch_size: int = 8, group=None):
        """model: ctpa_clip_b200 CTCLIP on a CUDA device; prompt_tokens: tokenised prompt pairs (2P, L) (`tokenize_prompts`)"""
        self.model, self.batch_size, self.group = model, batch_size, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        was_training = model.training
        model.eval()
        self.prompt_latents = model.prompt_latents(prompt_tokens)       # (2P, d), once per evaluation
        model.train(was_training)
        self.num_pathologies = self.prompt_latents.shape[0] // 2

    @torch.no_grad()
    def score(self, volumes: torch.Tensor) -> torch.Tensor:
        """(V, 1, f, h, w) CUDA volumes -> (V, P) prob[present]; each volume goes through the image tower once"""
        was_training = self.model.training
        self.model.eval()
        outs = [self.model.zero_shot_scores(None, volumes[i:i + self.batch_size], prompt_latents=self.prompt_latents)
                for i in range(0, volumes.shape[0], self.batch_size)]
        self.model.train(was_training)
        return torch.cat(outs)

    @torch.no_grad()
    def evaluate(self, dataset, labels=None):
        """dataset: indexable, item i -> volume (1, f, h, w) tensor (CPU or CUDA); this rank scores its round-robin shard.
        Returns (predicted_all (N, P) float32 ndarray in dataset order, real_all ndarray or None) on every rank."""
        n = len(dataset)
        mine = shard_indices(n, self.rank, self.world)
        dev = self.prompt_latents.device
        bs = self.batch_size
        batches = [mine[i:i + bs] for i in range(0, len(mine), bs)]
        rows = []
        if batches:
            # volumes go host -> device one by one into a double-buffered batch (no host-side stack: 221 MB per production
            # volume); the copy of batch i+1 runs on its own stream under the scoring of batch i. Pinned items copy by DMA.
            shape = tuple(torch.as_tensor(dataset[mine[0]]).shape)
            bufs = [torch.empty((bs, *shape), device=dev, dtype=torch.float32) for _ in range(2)]
            copy_stream = torch.cuda.Stream(device=dev)
            ready = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [torch.cuda.Event(), torch.cuda.Event()]
            for e in consumed:
                e.record()

            def prefetch(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i % 2])
                    for j, idx in enumerate(batches[i]):
                        bufs[i % 2][j].copy_(torch.as_tensor(dataset[idx]), non_blocking=True)
                    ready[i % 2].record(copy_stream)

            prefetch(0)
            for i, b in enumerate(batches):
                if i + 1 < len(batches):
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[i % 2])
                rows.append(self.score(bufs[i % 2][: len(b)]))
                consumed[i % 2].record()
        local = torch.cat(rows) if rows else torch.empty((0, self.num_pathologies), device=dev)
        pred = gather_rows(local, n, self.rank, self.world, self.group)
        real = None if labels is None else np.asarray(labels)
        return pred.cpu().numpy(), real


def aurocs(predicted_all: np.ndarray, real_all: np.ndarray, pathologies=PATHOLOGIES) -> dict:
    """per-pathology ROC AUC, the number `evaluate_internal` tabulates (evaluate.py:186-189 via sklearn roc_curve + auc);
    classes with a single label value are skipped as sklearn cannot score them."""
    from sklearn.metrics import auc, roc_curve
    out = {}
    for i, name in enumerate(pathologies[: predicted_all.shape[1]]):
        y = real_all[:, i]
        if y.min() == y.max():
            continue
        fpr, tpr, _ = roc_curve(y, predicted_all[:, i])
        out[name + "_auc"] = float(auc(fpr, tpr))
    return out
