"""CPU: the oracle restatement against the fixtures produced by the UNMODIFIED reference (tools/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import ctclip_oracle as O


def _probe(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def _run(name, training):
    fx = torch.load(f"tests/golden/ctclip_{name}.pt", weights_only=False)
    cfg = O.CONFIGS["mid" if name == "mid4" else name]
    sd = O.init_state_dict(cfg, fx["seed"])
    if training:
        sd = {k: v.requires_grad_(v.dtype.is_floating_point and v.numel() > 0 and "codebook" not in k and "beta" not in k)
              for k, v in sd.items()}
    txt = O.make_text_encoder(cfg, fx["seed"])
    video, ids, mask = O.make_inputs(cfg, fx["batch"], fx["seed"])
    return fx, sd, txt, O.ctclip_forward(sd, cfg, txt, ids, mask, video, training=training)


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_oracle_eval_matches_reference(name):
    fx, _, _, out = _run(name, False)
    with torch.no_grad():
        assert torch.equal(out["indices"].reshape(fx["batch"], -1).to(torch.int32), fx["indices"])
        assert torch.allclose(out["loss"], fx["loss_eval"], atol=1e-6)
        assert torch.allclose(out["text_latents"], fx["text_latents"], atol=1e-6)
        assert torch.allclose(out["image_latents"], fx["image_latents"], atol=1e-6)
        assert torch.allclose(out["pre_vq"].reshape(fx["batch"], -1, out["pre_vq"].shape[-1]), fx["pre_vq_full"], atol=1e-5)
        assert torch.allclose(out["sim"], fx["sim_eval"], atol=1e-5)


@pytest.mark.parametrize("name", ["tiny", "mid", "mid4"])
def test_oracle_gradients_match_reference(name):
    fx, sd, txt, out = _run(name, True)
    out["loss"].backward()
    assert torch.allclose(out["loss"].detach(), fx["loss_train"], atol=1e-6)
    text_params = dict(txt.named_parameters())
    checked = 0
    for k, g in fx["grads"].items():
        if k.startswith("text_transformer."):
            p = text_params.get(k[len("text_transformer."):])
            grad = p.grad if p is not None else None
        else:
            grad = sd[k].grad if k in sd else None
        if grad is None:
            continue
        assert abs(grad.norm().item() - g["norm"]) <= 1e-4 * max(1e-6, g["norm"]) + 1e-7, k
        proj = (grad * _probe(grad.shape, g["probe_seed"])).sum().item()
        assert abs(proj - g["proj"]) <= 2e-4 * max(abs(g["proj"]), g["norm"]) + 1e-6, k
        if g["full"] is not None:
            assert torch.allclose(grad, g["full"], atol=1e-5 * max(1.0, g["norm"])), k
        checked += 1
    assert checked > 90


def test_oracle_production_forward_matches_reference():
    """production config, B=2, ~10 s of CPU: indices bit-exact, latents / loss to fp32 round-off"""
    fx, _, _, out = _run("production", False)
    with torch.no_grad():
        agree = (out["indices"].reshape(2, -1).to(torch.int32) == fx["indices"]).float().mean().item()
        assert agree == 1.0
        assert torch.allclose(out["loss"], fx["loss_eval"], atol=1e-5)
        assert torch.allclose(out["image_latents"], fx["image_latents"], atol=1e-5)
        assert torch.allclose(out["pre_vq"].reshape(2, -1, 512)[:, :32], fx["pre_vq_head"], atol=1e-4)


def test_infonce_identity():
    """ct_clip.py:858-878 == 0.5 * (CE(L, arange) + CE(L^T, arange)) (SURVEY §4 property)"""
    torch.manual_seed(0)
    t = torch.nn.functional.normalize(torch.randn(5, 16), dim=-1)
    i = torch.nn.functional.normalize(torch.randn(5, 16), dim=-1)
    tau = torch.tensor(1.0)
    L = tau.exp() * t @ i.t()
    ce = torch.nn.functional.cross_entropy
    ref = 0.5 * (ce(L, torch.arange(5)) + ce(L.t(), torch.arange(5)))
    assert torch.allclose(O.clip_loss(t, i, tau), ref, atol=1e-6)
    assert float(O.clip_loss(t[:1], i[:1], tau)) == pytest.approx(0.0, abs=1e-6)  # B=1 -> 0


def test_temporal_peg_is_axis_permutation_on_cubic_grid():
    """attention.py:69-70 with ctvit.py:325-327: on a t=h=w grid the reshape equals a conv over (h, w, t), causal on h"""
    torch.manual_seed(0)
    b, n, d = 1, 4, 8
    sd = {"dsconv.weight": torch.randn(d, 1, 3, 3, 3), "dsconv.bias": torch.randn(d)}
    x = torch.randn(b, n, n, n, d)  # (b, t, h, w, d) canonical
    xin = x.permute(0, 2, 3, 1, 4).reshape(b * n * n, n, d)
    got = O.peg(sd, "", xin, (b, n, n, n)).reshape(b, n, n, n, d)  # indexed (h, w, t)
    y = x.permute(0, 2, 3, 1, 4).permute(0, 4, 1, 2, 3)  # (b, d, h, w, t): conv axes = (h, w, t)
    y = torch.nn.functional.pad(y, (1, 1, 1, 1, 2, 0))
    ref = torch.nn.functional.conv3d(y, sd["dsconv.weight"], sd["dsconv.bias"], groups=d).permute(0, 2, 3, 4, 1)
    assert torch.allclose(got, ref, atol=1e-5)


def test_vq_ema_matches_module_restatement():
    from oracle.vq_restatement import VectorQuantize
    torch.manual_seed(0)
    vq = VectorQuantize(dim=16, codebook_size=32).train()
    x = torch.randn(2, 40, 16)
    e0, c0 = vq._codebook.embed.clone(), vq._codebook.cluster_size.clone()
    q, idx, _ = vq(x)
    q2, idx2 = O.vq_assign(e0, x)
    assert torch.equal(idx, idx2)
    e1, c1 = O.vq_ema(e0, c0, x, idx2)
    assert torch.allclose(e1, vq._codebook.embed, atol=1e-6) and torch.allclose(c1, vq._codebook.cluster_size, atol=1e-6)
