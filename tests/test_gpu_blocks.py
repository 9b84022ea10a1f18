"""GPU (-m gpu): the block-level module surface — `Transformer.forward(x, video_shape, attn_bias)`, `PEG.forward(x, shape)`,
`Attention.forward(x, attn_bias=)`, `GEGLU` — called on their own with the reference's signatures and memory layouts
(attention.py:39-42,63-84,127-181,312-333; ctpa_report/vqa_meditron.py:107 reaches `.enc_spatial_transformer` this way),
against the CPU oracle restatement of the same blocks. bf16 tensor-core operands: 2.5e-2 of the tensor's max magnitude."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ctclip_oracle as O  # noqa: E402


def _vit(cfg, seed=0):
    from _common import build
    sd = O.init_state_dict(cfg, seed)
    m = build(cfg, sd, O.make_text_encoder(cfg, seed)).eval()
    return m.visual_transformer, sd


def _close(got, ref, tol):
    err = (got.float().cpu() - ref).abs().max().item()
    assert err <= tol * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_transformer_blocks_forward_in_the_reference_layouts(name):
    """spatial: '(b t) (h w) d' + the full (heads, n, n) bias tensor; temporal: '(b h w) t d', where the PEG reinterprets the
    flat buffer as (b, t, h, w, d) (a genuine scramble on the non-cubic tiny grid) — exactly the two calls of ctvit.py:315-329"""
    cfg = O.CONFIGS[name]
    vit, sd = _vit(cfg)
    b = 2
    t, h, w = O.grid_of(cfg)
    d = cfg["dim"]
    g = torch.Generator().manual_seed(3)
    tokens = torch.randn(b, t, h, w, d, generator=g)
    shape = (b, t, h, w)
    v = "visual_transformer."
    with torch.no_grad():
        bias = vit.spatial_rel_pos_bias(h, w)
        _close(bias, O.cpb_bias(sd, h, w), 1e-4)
        xs = tokens.reshape(b * t, h * w, d)
        got_s = vit.enc_spatial_transformer(xs.cuda(), attn_bias=bias, video_shape=shape)
        ref_s = O.transformer(sd, v + "enc_spatial_transformer.", xs, cfg["spatial_depth"], cfg["heads"], shape,
                              O.cpb_bias(sd, h, w))
        assert got_s.shape == ref_s.shape
        _close(got_s, ref_s, 2.5e-2)
        # an untagged copy of the bias (e.g. computed by someone else) is folded back into its table as well
        got_s2 = vit.enc_spatial_transformer(xs.cuda(), attn_bias=bias.clone(), video_shape=shape)
        assert torch.equal(got_s, got_s2)
        xt = ref_s.reshape(b, t, h, w, d).permute(0, 2, 3, 1, 4).reshape(b * h * w, t, d).contiguous()
        got_t = vit.enc_temporal_transformer(xt.cuda(), video_shape=shape)
        ref_t = O.transformer(sd, v + "enc_temporal_transformer.", xt, cfg["temporal_depth"], cfg["heads"], shape, None)
        _close(got_t, ref_t, 2.5e-2)


def test_peg_attention_geglu_modules_on_their_own():
    cfg = O.TINY
    vit, sd = _vit(cfg)
    b = 2
    t, h, w = O.grid_of(cfg)
    d = cfg["dim"]
    g = torch.Generator().manual_seed(5)
    p = "visual_transformer.enc_spatial_transformer.layers.0."
    peg, attn, _, ff = vit.enc_spatial_transformer.layers[0]
    with torch.no_grad():
        x5 = torch.randn(b, t, h, w, d, generator=g)
        _close(peg(x5.cuda()), O.peg(sd, p + "0.", x5, (b, t, h, w)), 1e-5)                    # (b, t, h, w, d) input
        x3 = x5.reshape(b * t, h * w, d)
        _close(peg(x3.cuda(), shape=(b, t, h, w)), O.peg(sd, p + "0.", x3, (b, t, h, w)), 1e-5)  # (b', n, d) + shape
        bias = vit.spatial_rel_pos_bias(h, w)
        _close(attn(x3.cuda(), attn_bias=bias), O.attention(sd, p + "1.", x3, cfg["heads"], O.cpb_bias(sd, h, w)), 2.5e-2)
        xt = torch.randn(7, 5, d, generator=g)                                                   # bias-free, n = 5
        _close(attn(xt.cuda()), O.attention(sd, p + "1.", xt, cfg["heads"], None), 2.5e-2)
        hin = torch.randn(11, 2 * 170, generator=g)
        a, gate = hin.chunk(2, dim=-1)
        _close(ff[2](hin.cuda()), torch.nn.functional.gelu(gate) * a, 1.5e-2)
    with pytest.raises(NotImplementedError):                                                     # not translation-invariant
        attn(x3.cuda(), attn_bias=torch.randn(cfg["heads"], h * w, h * w).cuda())
    with pytest.raises(NotImplementedError):                                                     # no autograd through a lone block
        peg(x5.cuda().requires_grad_())
