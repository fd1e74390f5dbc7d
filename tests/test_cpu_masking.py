"""CPU: univl_b200.masking — the MLM / MFM sampling rule of the reference dataloader (dataloader_howto100m.py:103-125,
:314-329) restated over whole tensors.  The rule is random, so the checks are the invariants the
reference's loop guarantees plus the selection statistics."""
import torch

from univl_b200 import masking


def _ids(n=400, W=48, seed=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(3, W + 1, (n,), generator=g)
    ids = torch.randint(1000, 30000, (n, W), generator=g)
    am = (torch.arange(W).unsqueeze(0) < lens.unsqueeze(1)).long()
    ids = ids * am
    ids[:, 0] = 101
    ids[torch.arange(n), lens - 1] = 102
    return ids, am, lens


def test_mask_tokens_invariants_and_statistics():
    ids, am, lens = _ids()
    g = torch.Generator().manual_seed(1)
    masked, labels = masking.mask_tokens(ids, am, generator=g)
    assert masked.dtype == torch.int64 and labels.dtype == torch.int64 and masked.shape == ids.shape
    n = ids.shape[0]
    # [CLS], [SEP] and padding are never touched / labelled; padded masked ids are 0 (dataloader_howto100m.py:108-110, :202)
    assert bool((labels[:, 0] == -1).all()) and bool((masked[:, 0] == 101).all())
    assert bool((labels[torch.arange(n), lens - 1] == -1).all()) and bool((masked[torch.arange(n), lens - 1] == 102).all())
    assert bool((labels[am == 0] == -1).all()) and bool((masked[am == 0] == 0).all())
    chosen = labels != -1
    assert bool((labels[chosen] == ids[chosen]).all())                 # label = the ORIGINAL id
    assert bool((masked[~chosen] == ids[~chosen]).all())               # unselected tokens are unchanged
    cand = (am == 1)
    cand[:, 0] = False
    cand[torch.arange(n), lens - 1] = False
    rate = chosen.sum().item() / cand.sum().item()
    assert abs(rate - 0.15) < 0.02, rate
    is_mask = (masked[chosen] == masking.MASK_ID).float().mean().item()
    kept = (masked[chosen] == ids[chosen]).float().mean().item()
    assert abs(is_mask - 0.8) < 0.04 and abs(kept - 0.1) < 0.03, (is_mask, kept)   # the rest: random vocabulary ids
    # deterministic under a seeded generator; 3-D [B, n_pair, W] inputs keep their shape
    m2, l2 = masking.mask_tokens(ids, am, generator=torch.Generator().manual_seed(1))
    assert torch.equal(m2, masked) and torch.equal(l2, labels)
    m3, l3 = masking.mask_tokens(ids.view(100, 4, 48), am.view(100, 4, 48), generator=torch.Generator().manual_seed(1))
    assert torch.equal(m3.view(400, 48), masked) and torch.equal(l3.view(400, 48), labels)


def test_mask_frames_invariants_and_statistics():
    g = torch.Generator().manual_seed(2)
    n, F, D = 300, 48, 16
    lens = torch.randint(1, F + 1, (n,), generator=g)
    vm = (torch.arange(F).unsqueeze(0) < lens.unsqueeze(1)).long()
    video = torch.randn(n, F, D, generator=g) + 3.0
    masked, labels = masking.mask_frames(video, vm, generator=g)
    chosen = labels != -1
    assert bool((labels[vm == 0] == -1).all())
    pos = torch.arange(F).expand(n, F)
    assert bool((labels[chosen] == pos[chosen]).all())                 # label = the frame's own index (:322-324)
    assert bool((masked[chosen] == 0).all()) and bool((masked[~chosen] == video[~chosen]).all())
    rate = chosen.sum().item() / vm.sum().item()
    assert abs(rate - 0.15) < 0.02, rate


def test_mask_pretrain_batch_keys_match_forward_signature():
    import inspect
    from univl_b200.modules.modeling import UniVL
    ids, am, _ = _ids(8)
    batch = {"input_ids": ids.view(8, 1, 48), "attention_mask": am.view(8, 1, 48),
             "token_type_ids": torch.zeros(8, 1, 48, dtype=torch.long),
             "video": torch.randn(8, 1, 48, 1024), "video_mask": torch.ones(8, 1, 48, dtype=torch.long)}
    out = masking.mask_pretrain_batch(batch, generator=torch.Generator().manual_seed(0))
    params = set(inspect.signature(UniVL.forward).parameters)
    assert set(out) <= params, set(out) - params
    for k in ("pairs_masked_text", "pairs_token_labels", "masked_video", "video_labels_index"):
        assert k in out
    assert out["pairs_masked_text"].shape == batch["input_ids"].shape
    assert out["masked_video"].shape == batch["video"].shape and out["video_labels_index"].shape == (8, 1, 48)
