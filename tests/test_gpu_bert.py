"""GPU parity of the native BERT text tower (ctpa_clip_b200/text/bert.py) against the reference's own text encoder:
transformers.BertModel in fp32 on the CPU (ct_clip.py:685-686 consumes its last_hidden_state). Tolerances are those of
bf16 tensor-core operands with fp32 accumulation: outputs within 2.5e-2 of the tensor maximum, parameter gradients within
5 % of each gradient's maximum."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bert(hidden, heads, inter, layers, max_pos, vocab=500, seed=0):
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    m = BertModel(BertConfig(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                             intermediate_size=inter, max_position_embeddings=max_pos, hidden_dropout_prob=0.0,
                             attention_probs_dropout_prob=0.0))
    return m


@pytest.mark.parametrize("hidden,heads,inter,layers,B,L,pad", [
    (64, 2, 128, 2, 3, 16, 8),      # tiny config of the test models: head dim 32, sequence shorter than one tile
    (256, 4, 512, 2, 2, 128, 40),   # head dim 64 (BERT-base), one full 128-row tile, ragged mask
    (128, 2, 256, 1, 2, 200, 0),    # sequence not a multiple of 64 / 128: every operand edge is TMA zero fill
    (192, 3, 384, 2, 2, 328, 150),  # fused attention path (head dim 64): ragged tiles, two whole key blocks of padding
])
def test_native_bert_forward_backward_matches_hf(hidden, heads, inter, layers, B, L, pad):
    from ctpa_clip_b200.text import NativeBert, supports
    from ctpa_clip_b200.text.bert import encode
    ref = _bert(hidden, heads, inter, layers, max(L, 32)).train()
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(1, 500, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.long)
    if pad:
        ids[:, L - pad:] = 0
        mask[:, L - pad:] = 0
        mask[0, L - pad: L - pad + 3] = 1   # ragged: different valid lengths inside the batch
    w = torch.randn(B, L, hidden, generator=g)
    out_ref = ref(ids, attention_mask=mask)[0]
    (out_ref * w).sum().backward()
    gref = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}

    import copy
    dev = copy.deepcopy(ref).cuda()
    for p in dev.parameters():
        p.grad = None
    assert supports(dev)
    eng = NativeBert(dev)
    out = encode(eng, ids.cuda(), mask.cuda(), True)
    assert out.shape == out_ref.shape
    err = (out.cpu() - out_ref.detach()).abs().max().item()
    assert err < 2.5e-2 * out_ref.abs().max().item(), f"forward max err {err}"
    (out * w.cuda()).sum().backward()
    # key.bias has an exactly-zero true gradient (softmax is invariant to a per-query constant): its reference value is
    # rounding noise, so every tensor is also given an absolute floor of 2e-3 of the largest bias gradient
    floor = 2e-3 * max(g.abs().max().item() for n, g in gref.items() if n.endswith(".bias"))
    worst = ("", 0.0)
    for n, p in dev.named_parameters():
        if n.startswith("pooler."):
            continue
        assert p.grad is not None, n
        ref_g = gref[n]
        rel = (p.grad.cpu() - ref_g).abs().max().item() / (ref_g.abs().max().item() + floor)
        if rel > worst[1]:
            worst = (n, rel)
    assert worst[1] < 5e-2, f"gradient of {worst[0]} off by {worst[1]:.3f} of its maximum"


def test_native_bert_dropout_is_consistent_between_forward_and_backward():
    """train-mode dropout: masks are regenerated from (seed, index) in the backward; check d(sum out)/d(out-bias of the last
    layer) analytically (= kept fraction / (1-p) per column through LayerNorm is hard to state) -> instead check that
    the gradient is finite, differs from the p=0 gradient, and that ~p of the attention probabilities are dropped."""
    from ctpa_clip_b200 import ops
    S = torch.randn(2 * 2, 64, 64, device="cuda")
    mask = torch.ones(2, 64, dtype=torch.long, device="cuda")
    P, Pd = ops.bert_softmax_fwd(S, mask, 2, 2, 64, 0.125, 0.1, 1234)
    dropped = (Pd == 0).float().mean().item()
    assert 0.07 < dropped < 0.13
    kept = Pd != 0
    assert torch.allclose(Pd[kept].float(), (P[kept].float() / 0.9), rtol=2e-2, atol=1e-4)
    # backward regenerates the same mask: gradient is exactly zero where the probability was dropped
    dP = torch.ones_like(S)
    dS0 = ops.bert_softmax_bwd(P, dP, 2, 2, 64, 0.125, 0.0, 1234)
    assert dS0.float().abs().max().item() < 1e-2          # softmax gradient of a constant upstream is ~0
    dS = ops.bert_softmax_bwd(P, dP, 2, 2, 64, 0.125, 0.1, 1234)
    assert torch.isfinite(dS.float()).all() and dS.float().abs().max().item() > 1e-3
    y = torch.randn(1024, 64, device="cuda")
    a = ops.dropout_add(y, None, 0.1, 77)
    b = ops.dropout_add(y, None, 0.1, 77)
    assert torch.equal(a, b) and 0.07 < (a == 0).float().mean().item() < 0.13


@pytest.mark.parametrize("L,pad,p_drop", [(128, 0, 0.0), (200, 70, 0.0), (512, 256, 0.1), (96, 10, 0.25)])
def test_fused_attention_equals_gemm_softmax_path(L, pad, p_drop):
    """csrc/bert_attn.cu (scores in registers, mma.sync) against the batched tcgen05 GEMM + softmax kernels on the same packed
    QKV: same math, same stateless dropout mask (element index and seed), bf16 operands in both"""
    import math
    from ctpa_clip_b200 import ops
    B, H, D = 2, 3, 192
    g = torch.Generator(device="cuda").manual_seed(L + pad)
    qkv = (torch.randn(B * L, 3 * D, device="cuda", generator=g) * 1.5).bfloat16()
    dout = torch.randn(B * L, D, device="cuda", generator=g).bfloat16()
    mask = torch.ones(B, L, dtype=torch.long, device="cuda")
    if pad:
        mask[:, L - pad:] = 0
        mask[1, L - pad: L - pad + 5] = 1
    scale, seed = 1.0 / math.sqrt(64), 4242

    class M:   # the unfused product helpers of NativeBert only need these attributes
        pass
    from ctpa_clip_b200.text.bert import NativeBert
    eng = NativeBert.__new__(NativeBert)
    eng.D, eng.H, eng.hd = D, H, 64
    S = eng._scores(qkv, B, L)
    P, Pd = ops.bert_softmax_fwd(S, mask, B, H, L, scale, p_drop, seed)
    cv_ref = eng._pv(Pd if Pd is not None else P, qkv, B, L)
    dq_ref = eng._attention_backward_unfused(qkv, P, Pd, dout, B, L, scale, p_drop, seed - 1, qkv.device)

    cv, lse = ops.bert_attn_fwd(qkv, mask, B, H, L, D, p_drop, seed)
    dqkv = ops.bert_attn_bwd(qkv, mask, cv, lse, dout, B, H, L, D, p_drop, seed)
    def close(a, b, tol):
        err = (a.float() - b.float()).abs().max().item()
        assert err <= tol * b.float().abs().max().item(), f"{err} vs {b.float().abs().max().item()}"
    close(cv, cv_ref, 2e-2)
    for j, name in enumerate(("dq", "dk", "dv")):
        close(dqkv[:, j * D:(j + 1) * D], dq_ref[:, j * D:(j + 1) * D], 3e-2)
    # lse against the fp32 definition (log2 domain), valid queries only
    ref = torch.logsumexp(S.float() * scale + torch.where(mask[:, None, None, :] != 0, 0.0, float("-inf")).expand(B, H, L, L)
                          .reshape(B * H, L, L), dim=-1) / math.log(2.0)
    assert (lse - ref).abs().max().item() < 2e-2
