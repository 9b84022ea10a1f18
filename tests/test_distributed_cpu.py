"""CPU, gloo, world_size 2: the global-batch InfoNCE partitioning — each rank differentiates its own rows
of the all-gathered logits and the SUM of rank gradients equals the single-process gradient on the concatenated batch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ctclip_oracle as O


def _local_grads(T, I, tau, row0, rows):
    """restates csrc/loss.cu clip_grad_kernel in torch: gradient w.r.t. rows [row0, row0+rows) and the dtau share"""
    B = T.shape[0]
    L = tau.exp() * T @ I.t()
    row_lse, col_lse = torch.logsumexp(L, dim=1), torch.logsumexp(L, dim=0)
    G = (torch.exp(L - row_lse[:, None]) + torch.exp(L - col_lse[None, :]) - 2 * torch.eye(B)) / (2 * B)
    sl = slice(row0, row0 + rows)
    dT = tau.exp() * G[sl] @ I
    dI = tau.exp() * G[:, sl].t() @ T
    dtau = (G[sl] * L[sl]).sum()
    loss = (row_lse + col_lse - 2 * torch.diagonal(L)).sum() / (2 * B)
    return loss, dT, dI, dtau


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    b, d = 3, 16
    T_all = torch.nn.functional.normalize(torch.randn(world * b, d), dim=-1)
    I_all = torch.nn.functional.normalize(torch.randn(world * b, d), dim=-1)
    tau = torch.tensor(0.7)
    local = torch.stack((T_all[rank * b:(rank + 1) * b], I_all[rank * b:(rank + 1) * b]))
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    g = torch.stack(gathered)
    T, I = g[:, 0].reshape(world * b, d), g[:, 1].reshape(world * b, d)
    loss, dT, dI, dtau = _local_grads(T, I, tau, rank * b, b)
    # a parameter shared by all ranks (here: tau) receives the SUM of the rank shares
    dist.all_reduce(dtau)
    ret[rank] = (loss, dT, dI, dtau)
    dist.destroy_process_group()


def test_global_infonce_partition_two_ranks():
    world, b, d = 2, 3, 16
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29617, ret), nprocs=world, join=True)
    torch.manual_seed(0)
    T = torch.nn.functional.normalize(torch.randn(world * b, d), dim=-1).requires_grad_()
    I = torch.nn.functional.normalize(torch.randn(world * b, d), dim=-1).requires_grad_()
    tau = torch.tensor(0.7, requires_grad=True)
    loss = O.clip_loss(T, I, tau)           # single process on the concatenated batch (the parity oracle for W>1)
    loss.backward()
    for r in range(world):
        l, dT, dI, dtau = ret[r]
        assert torch.allclose(l, loss.detach(), atol=1e-6)
        assert torch.allclose(dT, T.grad[r * b:(r + 1) * b], atol=1e-6)
        assert torch.allclose(dI, I.grad[r * b:(r + 1) * b], atol=1e-6)
        assert torch.allclose(dtau, tau.grad, atol=1e-6)


# ---------------------------------------------------------------- zero-shot evaluation: replicas, round-robin shards (config 5)
def _gather_worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ctpa_clip_b200.inference import gather_rows, shard_indices
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)         # row i = "scores of volume i"
    mine = shard_indices(n, rank, world)
    out = gather_rows(full[mine], n, rank, world)
    ret[rank] = out
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8, 1])
def test_zero_shot_shard_and_gather_two_ranks(n):
    """every rank ends with the (N, P) score matrix in dataset order, also when N is not a multiple of the world size"""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_gather_worker, args=(world, 29618 + n, n, ret), nprocs=world, join=True)
    want = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    for r in range(world):
        assert torch.equal(ret[r], want)


def test_zero_shot_prompts_and_aurocs():
    from ctpa_clip_b200 import inference as inf
    texts = inf.prompt_texts()
    assert len(texts) == 36 and texts[0] == "Medical material is present." and texts[1] == "Medical material is not present."
    assert texts[22:24] == ["Pulmonary Embolism is present.", "Pulmonary Embolism is not present."]   # ctclip_inference.py:318
    pred = torch.tensor([[0.9, 0.2], [0.8, 0.7], [0.3, 0.6], [0.1, 0.4]]).numpy()
    real = torch.tensor([[1, 0], [1, 1], [0, 0], [0, 1]]).numpy()
    a = inf.aurocs(pred, real, ["a", "b"])
    assert a["a_auc"] == 1.0 and abs(a["b_auc"] - 0.75) < 1e-9


# ---------------------------------------------------------------- factor gather of the rank-B to_visual_latent gradient
def _factor_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7 + rank)
    b, n, k = 3, 8, 40
    dy, x = torch.randn(b, n, generator=g), torch.randn(b, k, generator=g)
    # route 1 (reference DDP semantics, summed): all-reduce of the N x K product
    prod = dy.t() @ x
    dist.all_reduce(prod)
    # route 2 (ct_clip.LinearFunction.backward, factor_gather): all-gather both factors, multiply locally
    dy_all, x_all = torch.empty(world * b, n), torch.empty(world * b, k)
    dist.all_gather_into_tensor(dy_all, dy)
    dist.all_gather_into_tensor(x_all, x)
    ret[rank] = (prod, dy_all.t() @ x_all)
    dist.destroy_process_group()


def test_factor_gather_equals_allreduce_of_product():
    """dW = sum_r dL_r^T E_r = [dL_0; dL_1]^T [E_0; E_1]: gathering 2*B_loc*(N+K) numbers replaces reducing N*K"""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_factor_worker, args=(world, 29631, ret), nprocs=world, join=True)
    for r in range(world):
        prod, fact = ret[r]
        assert torch.allclose(prod, fact, atol=1e-5)
    assert torch.equal(ret[0][1], ret[1][1])        # every rank holds the identical global-batch gradient
