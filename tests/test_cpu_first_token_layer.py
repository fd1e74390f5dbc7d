"""CPU: the restructuring behind ops.EncoderLayerClsFn, checked at the oracle level in fp32.

Under `_cross_similarity` (reference modules/modeling.py:341-375) only token 0 of the last cross layer reaches the loss
(pooler -> similarity_dense).  The product therefore runs the last layer's query side on the first-token rows only (the
K/V projections stay dense).  Here the same restructuring is written with the oracle's primitives and compared with
the oracle's dense layer: identical pooled logits and identical gradients for every parameter and for the layer input
(the rows the dense form computes and discards receive exactly zero gradient)."""
import torch

from oracle import synth
from oracle import univl_oracle as O


def _last_layer_first_token(x, add_mask, sd, pfx):
    """first-token rows of O.encoder_layer(x, ...): query side on x[:, :1], keys / values on all of x"""
    xq = x[:, :1]
    ctx = O.multi_head_attention(xq, x, add_mask, sd, pfx + "attention.self.")
    att = O.dense_residual_norm(ctx, xq, sd, pfx + "attention.output.")
    inter = O.gelu(O.linear(att, sd, pfx + "intermediate.dense"))
    return O.dense_residual_norm(inter, att, sd, pfx + "output.")


def _logits(first_token_only, sd, x, mask, n_layers):
    add = O.additive_mask(mask, x.dtype)
    h = x
    for n in range(n_layers - 1):
        h = O.encoder_layer(h, add, sd, "cross.encoder.layer.%d." % n)
    pfx = "cross.encoder.layer.%d." % (n_layers - 1)
    last = _last_layer_first_token(h, add, sd, pfx) if first_token_only else O.encoder_layer(h, add, sd, pfx)
    pooled = O.pooler(last, sd, "cross.pooler.")
    return O.linear(pooled, sd, "similarity_dense").squeeze(-1)


def test_first_token_only_last_layer_equals_dense_layer():
    cfg = synth.task_config(mode="ft_align", batch_size=3, text_layers=1, visual_layers=1, cross_layers=2,
                            max_words=6, max_frames=5)
    base = synth.make_state_dict(cfg, seed=3)
    keys = [k for k in base if k.startswith("cross.encoder.") or k.startswith("cross.pooler.")
            or k.startswith("similarity_dense")]
    g = torch.Generator().manual_seed(11)
    n_seq, S, H = 5, 11, 768
    x0 = torch.randn(n_seq, S, H, generator=g)
    mask = (torch.arange(S).unsqueeze(0) < torch.tensor([11, 7, 1, 9, 4]).unsqueeze(1)).long()
    results = []
    for mode in (False, True):
        sd = {k: base[k].clone().double().requires_grad_(True) for k in keys}
        x = x0.clone().double().requires_grad_(True)
        logits = _logits(mode, sd, x, mask, 2)
        (logits * torch.linspace(-1.0, 1.0, n_seq, dtype=torch.float64)).sum().backward()
        results.append((logits.detach(), x.grad, {k: v.grad for k, v in sd.items()}))
    (l0, gx0, gp0), (l1, gx1, gp1) = results
    assert torch.allclose(l0, l1, rtol=0, atol=1e-12)
    assert torch.allclose(gx0, gx1, rtol=0, atol=1e-12)
    for k in keys:
        a, b = gp0[k], gp1[k]
        assert (a is None) == (b is None), k
        if a is not None:
            assert torch.allclose(a, b, rtol=0, atol=1e-11), k
    # and the dense form really does give the discarded rows a zero upstream gradient: d(loss)/d(last-layer output)
    sd = {k: base[k].clone().double() for k in keys}
    add = O.additive_mask(mask, torch.float64)
    h = O.encoder_layer(x0.double(), add, sd, "cross.encoder.layer.0.")
    last = O.encoder_layer(h, add, sd, "cross.encoder.layer.1.").requires_grad_(True)
    O.linear(O.pooler(last, sd, "cross.pooler."), sd, "similarity_dense").sum().backward()
    assert float(last.grad[:, 1:].abs().max()) == 0.0 and float(last.grad[:, 0].abs().max()) > 0.0
