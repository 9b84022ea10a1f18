"""GPU (-m gpu), ONE device: the fused latent exchange + global-batch InfoNCE over symmetric buffers
(`ctclip_clip_loss_allgather`, csrc/symm.cu) with all ranks emulated in ONE cooperative launch
(`ctclip_clip_loss_allgather_emulated`: blockIdx.z plays the rank, every rank's buffer lives on this device) against the
single-device `ctclip_clip_loss` on the concatenated batch — loss and latent gradients bit-exact (same fp32 summation order; the
temperature share is an atomic sum, compared to 1e-5), several steps in a row so that both buffer parities and the
step-valued flags are exercised. Mutually waiting kernels are never separate launches on one GPU (B200_PROFILING.md); the
true multi-process test runs one rank per device in tests/test_gpu_multi.py."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _buffers(world, b, d):
    from ctpa_clip_b200 import _lib
    lib = _lib.lib()
    lib.ctclip_symm_latent_bytes.restype = C.c_size_t
    nbytes = lib.ctclip_symm_latent_bytes(b, d, world)
    assert nbytes > 0
    # plain torch allocations are fine here: nothing is exported through CUDA IPC
    keep = [torch.zeros(nbytes // 4, device="cuda", dtype=torch.float32) for _ in range(world)]
    table = (C.c_void_p * world)(*[t.data_ptr() for t in keep])
    return keep, table


@pytest.mark.parametrize("world,b,d,steps", [(2, 8, 512, 5), (4, 3, 64, 5), (8, 8, 512, 3)])
def test_emulated_exchange_matches_single_device(world, b, d, steps):
    from ctpa_clip_b200 import _lib, ops
    keep, table = _buffers(world, b, d)
    status = torch.zeros(1, device="cuda", dtype=torch.int32)
    n0 = _lib.launch_count()
    for step in range(steps):
        g = torch.Generator().manual_seed(100 + step)
        T_all = torch.nn.functional.normalize(torch.randn(world * b, d, generator=g), dim=-1).cuda()
        I_all = torch.nn.functional.normalize(torch.randn(world * b, d, generator=g), dim=-1).cuda()
        tau = torch.tensor([0.3 + 0.2 * step], device="cuda")
        loss, dT, dI, dtau = ops.clip_loss_allgather_emulated(T_all, I_all, tau, world, table, step + 1, status)
        torch.cuda.synchronize()
        for r in range(world):
            want = ops.clip_loss(T_all, I_all, tau, r * b, b, want_grad=True)   # single-device kernel on the full batch
            sl = slice(r * b, (r + 1) * b)
            assert torch.equal(loss[r], want[0]), (step, r, float(loss[r]), float(want[0]))
            assert torch.equal(dT[sl], want[1]) and torch.equal(dI[sl], want[2]), (step, r)
            assert torch.allclose(dtau[r], want[3], rtol=1e-5, atol=1e-7), (step, r)
    assert int(status.item()) == 0
    assert _lib.launch_count() - n0 > steps


def test_missing_peer_times_out_raises_status_and_optimiser_skips_the_update():
    """a rank whose peer never publishes its flag: the loss is NaN, *status is raised (not a hang), and ctclip_adam_step leaves
    parameters / moments / gradients untouched when the gradient norm is not finite (a timeout must never become a
    training update)"""
    from ctpa_clip_b200 import ops, symm
    world, b, d = 2, 4, 64
    keep, table = _buffers(world, b, d)
    status = torch.zeros(1, device="cuda", dtype=torch.int32)
    T = torch.nn.functional.normalize(torch.randn(b, d), dim=-1).cuda()
    I = torch.nn.functional.normalize(torch.randn(b, d), dim=-1).cuda()
    tau = torch.tensor([0.5], device="cuda")
    symm.set_timeout(0.2)
    try:                                                              # rank 0 alone: rank 1 never arrives
        loss, dT, dI, dtau = ops.clip_loss_allgather(T, I, tau, 0, world, table, 1, status)
        torch.cuda.synchronize()
    finally:
        symm.set_timeout(600.0)
    assert int(status.item()) == 1 and torch.isnan(loss)
    n = 1024
    p, m, v = torch.randn(n, device="cuda"), torch.rand(n, device="cuda"), torch.rand(n, device="cuda")
    gbad = torch.full((n,), float("nan"), device="cuda")
    p0, m0, v0 = p.clone(), m.clone(), v.clone()
    ns = torch.zeros(1, device="cuda")
    skipped = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.sumsq(gbad, ns)
    ops.adam_step(p, gbad, m, v, None, 1e-3, 0.9, 0.99, 1e-8, 1, norm_sq=ns, max_norm=0.5, skipped=skipped)
    torch.cuda.synchronize()
    assert int(skipped.item()) == 1
    assert torch.equal(p, p0) and torch.equal(m, m0) and torch.equal(v, v0) and torch.isnan(gbad).all()
