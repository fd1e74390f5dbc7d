"""GPU: KV-cached caption decoding (univl_b200/caption.py) against the full-prefix path the reference drives
(`UniVL.decoder_caption`, main_task_caption.py:434-477): per-step last-token logits under teacher forcing, and the
beam search end to end."""
import pytest
import torch

from oracle import synth
from tests.model_util import build_model, to_device

pytestmark = pytest.mark.gpu


def _setup(batch_size=3, W=16, F=16):
    cfg = synth.task_config(mode="caption", batch_size=batch_size, text_layers=2, visual_layers=1, cross_layers=1,
                            decoder_layers=2, max_words=W, max_frames=F)
    batch = to_device(synth.make_batch(cfg, seed=31))
    model = build_model(cfg, seed=0)
    model.eval()
    with torch.no_grad():
        seq, vis = model.get_sequence_visual_output(batch["input_ids"], batch["token_type_ids"],
                                                    batch["attention_mask"], batch["video"], batch["video_mask"])
    flat = {k: v.view(-1, *v.shape[2:]) for k, v in batch.items()}
    return cfg, model, flat, seq, vis


def test_cached_decoder_matches_full_prefix_logits():
    from univl_b200.caption import CachedCaptionDecoder
    cfg, model, b, seq, vis = _setup()
    n, n_beam, L = seq.shape[0], 2, 9
    g = torch.Generator().manual_seed(3)
    tokens = torch.randint(1000, 30522, (n * n_beam, L), generator=g).cuda()
    tokens[:, 0] = 101
    dec = CachedCaptionDecoder(model, seq, vis, b["attention_mask"], b["video_mask"], n_beam, cfg.max_words)
    rep = lambda t: t.repeat_interleave(n_beam, 0)   # what the reference's `repeat(1, n_bm, 1).view(...)` builds
    worst = 0.0
    with torch.no_grad():
        for t in range(L):
            got = dec.step(tokens[:, t].contiguous())
            prefix = tokens[:, :t + 1].contiguous()
            want = model.decoder_caption(rep(seq), rep(vis), rep(b["input_ids"]), rep(b["attention_mask"]),
                                         rep(b["video_mask"]), prefix, torch.ones_like(prefix), shaped=True,
                                         get_logits=True)[:, -1]
            err = float((got - want).abs().max())
            scale = float(want.abs().max())
            worst = max(worst, err / max(scale, 1.0))
            assert err <= 2.0 ** -6 * max(scale, 1.0) + 2e-2, (t, err, scale)
    # beam reordering / instance removal keep the cache rows consistent
    with torch.no_grad():
        origin = torch.tensor([[1, 0], [0, 0], [1, 1]], device="cuda")
        dec.reorder(origin)
        rows = (torch.arange(n, device="cuda").unsqueeze(1) * n_beam + origin).reshape(-1)
        tokens = tokens.index_select(0, rows)
        dec.select([0, 2])
        keep_rows = torch.tensor([0, 1, 4, 5], device="cuda")
        tokens = tokens.index_select(0, keep_rows)
        nxt = torch.randint(1000, 30522, (4,), generator=g).cuda()
        got = dec.step(nxt)
        prefix = torch.cat([tokens, nxt.unsqueeze(1)], 1)
        sel = torch.tensor([0, 2], device="cuda")
        rs = lambda t: rep(t.index_select(0, sel))
        want = model.decoder_caption(rs(seq), rs(vis), rs(b["input_ids"]), rs(b["attention_mask"]), rs(b["video_mask"]),
                                     prefix, torch.ones_like(prefix), shaped=True, get_logits=True)[:, -1]
        assert float((got - want).abs().max()) <= 2.0 ** -6 * max(float(want.abs().max()), 1.0) + 2e-2


def test_beam_search_runs_and_scores_match_full_prefix_search():
    from univl_b200.caption import beam_search
    cfg, model, b, seq, vis = _setup(batch_size=3, W=16, F=16)
    n_beam, max_words = 3, 6
    hyps, scores = beam_search(model, seq, vis, b["attention_mask"], b["video_mask"], max_words, n_beam=n_beam)
    assert len(hyps) == seq.shape[0] and all(1 <= len(h) <= max_words for h in hyps)
    # score of the returned hypothesis under the full-prefix decoder (teacher forcing) equals the beam score
    rep = lambda t: t
    with torch.no_grad():
        for i, h in enumerate(hyps):
            ids = torch.tensor([[101] + h[:-1]], device="cuda")
            one = lambda t: t[i:i + 1]
            logits = model.decoder_caption(one(seq), one(vis), one(b["input_ids"]), one(b["attention_mask"]),
                                           one(b["video_mask"]), ids, torch.ones_like(ids), shaped=True,
                                           get_logits=True)[0]
            lp = torch.log_softmax(logits, -1)
            total = float(sum(lp[t, tok] for t, tok in enumerate(h)))
            assert abs(total - scores[i]) <= 0.05 * len(h) + 0.05, (i, total, scores[i])
            # and no single-token alternative at the first step beats the chosen beam set's best first token by more
            # than the tolerance (sanity of the top-k)
            assert float(lp[0].max()) >= float(lp[0, h[0]]) - 1e-6
