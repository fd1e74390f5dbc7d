"""Device-side MLM / MFM masking for pretraining batches (SURVEY.md §8f#4 "dataloaders -> device-side masking").

The reference samples the 15 % masks token by token in the CPU dataloader (dataloaders/dataloader_howto100m.py:103-125
for words, :314-329 for frames) and ships `pairs_masked_text`, `pairs_token_labels`, `masked_video`,
`video_labels_index` to the GPU as four extra tensors.  Here the same sampling rule runs on the tensors that are
already on the device (torch ops only: this is input preparation, not part of the kernels' hot path), so a pretraining
input pipeline needs to copy only `input_ids`, `attention_mask` and `video`.  The rule, restated:

  words  : position 0 ([CLS]) and the last real position ([SEP]) are never candidates.  Every other real token is
           selected with probability 0.15; a selected token becomes [MASK] with probability 0.8, a uniformly random
           vocabulary id with probability 0.1, stays itself with probability 0.1; its label is the ORIGINAL id.  All other
           positions (and padding) carry label -1; padded positions of the masked ids are 0.
  frames : every real frame is selected with probability 0.15 and replaced by zeros; its label is its own index j,
           every other position (and padding) -1.

The outputs have the shapes, dtypes and conventions `UniVL.forward(..., pairs_masked_text, pairs_token_labels,
masked_video, video_labels_index)` expects (reference modules/modeling.py:188-190).
"""
import torch

MASK_ID = 103        # "[MASK]" in bert-base-uncased's vocab.txt
VOCAB_SIZE = 30522


def mask_tokens(input_ids, attention_mask, p=0.15, mask_id=MASK_ID, vocab_size=VOCAB_SIZE, generator=None):
    """input_ids / attention_mask: integer [..., W] on any device -> (masked_ids, token_labels), both int64 [..., W]."""
    if input_ids.shape != attention_mask.shape:
        raise ValueError("mask_tokens: input_ids %s and attention_mask %s differ in shape"
                         % (tuple(input_ids.shape), tuple(attention_mask.shape)))
    ids = input_ids.long()
    real = attention_mask.bool()
    W = ids.shape[-1]
    pos = torch.arange(W, device=ids.device).expand(ids.shape)
    last = real.long().sum(-1, keepdim=True) - 1                     # index of [SEP]
    candidate = real & (pos != 0) & (pos != last)
    u = torch.rand(ids.shape, device=ids.device, generator=generator)
    chosen = candidate & (u < p)
    kind = u / p                                                      # the reference re-uses the same draw (prob /= 0.15)
    random_ids = torch.randint(0, vocab_size, ids.shape, device=ids.device, generator=generator)
    masked = torch.where(chosen & (kind < 0.8), torch.full_like(ids, mask_id), ids)
    masked = torch.where(chosen & (kind >= 0.8) & (kind < 0.9), random_ids, masked)
    masked = torch.where(real, masked, torch.zeros_like(ids))
    labels = torch.where(chosen, ids, torch.full_like(ids, -1))
    return masked, labels


def mask_frames(video, video_mask, p=0.15, generator=None):
    """video: float [..., F, D]; video_mask: integer [..., F] -> (masked_video like video, video_labels_index int64 [..., F])."""
    if video.shape[:-1] != video_mask.shape:
        raise ValueError("mask_frames: video %s and video_mask %s do not match"
                         % (tuple(video.shape), tuple(video_mask.shape)))
    real = video_mask.bool()
    F = video_mask.shape[-1]
    u = torch.rand(video_mask.shape, device=video.device, generator=generator)
    chosen = real & (u < p)
    masked = torch.where(chosen.unsqueeze(-1), torch.zeros_like(video), video)
    pos = torch.arange(F, device=video.device).expand(video_mask.shape)
    labels = torch.where(chosen, pos, torch.full_like(pos, -1))
    return masked, labels


def mask_pretrain_batch(batch, p=0.15, generator=None):
    """add `pairs_masked_text`, `pairs_token_labels`, `masked_video`, `video_labels_index` to a batch dict keyed by
    UniVL.forward's argument names (input_ids, attention_mask, video, video_mask), sampled on the tensors' own device."""
    out = dict(batch)
    out["pairs_masked_text"], out["pairs_token_labels"] = mask_tokens(batch["input_ids"], batch["attention_mask"], p,
                                                                      generator=generator)
    out["masked_video"], out["video_labels_index"] = mask_frames(batch["video"], batch["video_mask"], p,
                                                                 generator=generator)
    return out
