"""Deterministic synthetic weights and batches shared by the oracle, the golden generator, the tests and bench.py.

TEST / MEASUREMENT INFRASTRUCTURE — not part of the product path.

Nothing here depends on the reference or on univl_b200: the key/shape table below restates the checkpoint layout
contract (reference modules/modeling.py:134-170, module_bert.py:118-330,
module_visual.py:104-311, module_cross.py:109-290, module_decoder.py:195-349) so that the same `state_dict` can be
loaded into the unmodified reference (in the authoring container, by oracle/make_golden.py), into the oracle and
into the CUDA model.  torch's CPU generator is bit-reproducible across hosts for a fixed torch version, which
is what lets the golden fixtures travel without the 600 MB of weights.
"""
import argparse

import torch

H = 768
I = 3072
VOCAB = 30522
VIDEO_DIM = 1024


def task_config(mode="ft_joint", batch_size=4, n_gpu=1, n_pair=1, max_words=48, max_frames=48, text_layers=12,
                visual_layers=6, cross_layers=2, decoder_layers=3, use_mil=False, margin=0.1, **extra):
    """argparse.Namespace with the fields UniVL.__init__ reads (reference modules/modeling.py:110-184).

    mode: ft_joint | ft_align | caption | pretrain1 | pretrain2
    """
    ns = argparse.Namespace(
        do_pretrain=mode.startswith("pretrain"), do_train=True, task_type=None, stage_two=False,
        train_sim_after_cross=False, batch_size=batch_size, n_gpu=n_gpu, n_pair=n_pair, margin=margin,
        negative_weighting=1, hard_negative_rate=0.5, use_mil=use_mil, video_dim=VIDEO_DIM, max_words=max_words,
        max_frames=max_frames, local_rank=0, text_num_hidden_layers=text_layers,
        visual_num_hidden_layers=visual_layers, cross_num_hidden_layers=cross_layers,
        decoder_num_hidden_layers=decoder_layers, mode=mode)
    if mode == "ft_joint":
        ns.task_type = "retrieval"
    elif mode == "ft_align":
        ns.task_type = "retrieval"
        ns.train_sim_after_cross = True
    elif mode == "caption":
        ns.task_type = "caption"
        ns.stage_two = True
    elif mode == "pretrain1":
        ns.use_mil = True
    elif mode == "pretrain2":
        ns.stage_two = True
        ns.use_mil = True
    else:
        raise ValueError(mode)
    for k, v in extra.items():
        setattr(ns, k, v)
    return ns


def _encoder_layer_spec(pfx):
    out = []
    for n in ("query", "key", "value"):
        out += [(pfx + "attention.self.%s.weight" % n, (H, H)), (pfx + "attention.self.%s.bias" % n, (H,))]
    out += [(pfx + "attention.output.dense.weight", (H, H)), (pfx + "attention.output.dense.bias", (H,)),
            (pfx + "attention.output.LayerNorm.weight", (H,)), (pfx + "attention.output.LayerNorm.bias", (H,)),
            (pfx + "intermediate.dense.weight", (I, H)), (pfx + "intermediate.dense.bias", (I,)),
            (pfx + "output.dense.weight", (H, I)), (pfx + "output.dense.bias", (H,)),
            (pfx + "output.LayerNorm.weight", (H,)), (pfx + "output.LayerNorm.bias", (H,))]
    return out


def _decoder_layer_spec(pfx):
    out = []
    for att in ("slf_attn", "enc_attn"):
        for n in ("query", "key", "value"):
            out += [(pfx + "%s.att.%s.weight" % (att, n), (H, H)), (pfx + "%s.att.%s.bias" % (att, n), (H,))]
        out += [(pfx + att + ".output.dense.weight", (H, H)), (pfx + att + ".output.dense.bias", (H,)),
                (pfx + att + ".output.LayerNorm.weight", (H,)), (pfx + att + ".output.LayerNorm.bias", (H,))]
    out += [(pfx + "intermediate.dense.weight", (I, H)), (pfx + "intermediate.dense.bias", (I,)),
            (pfx + "output.dense.weight", (H, I)), (pfx + "output.dense.bias", (H,)),
            (pfx + "output.LayerNorm.weight", (H,)), (pfx + "output.LayerNorm.bias", (H,))]
    return out


def _mlm_head_spec(pfx, out_bias_dim):
    return [(pfx + "predictions.bias", (out_bias_dim,)),
            (pfx + "predictions.transform.dense.weight", (H, H)), (pfx + "predictions.transform.dense.bias", (H,)),
            (pfx + "predictions.transform.LayerNorm.weight", (H,)),
            (pfx + "predictions.transform.LayerNorm.bias", (H,))]


# alias key -> owner key (same storage in the reference)
def tied_keys(cfg):
    ties = {}
    has_cross = cfg.stage_two or cfg.train_sim_after_cross
    has_decoder = has_cross and not cfg.train_sim_after_cross
    if has_decoder:
        ties["decoder.embeddings.word_embeddings.weight"] = "bert.embeddings.word_embeddings.weight"
        ties["decoder.embeddings.position_embeddings.weight"] = "bert.embeddings.position_embeddings.weight"
        ties["decoder.classifier.cls.predictions.decoder.weight"] = "bert.embeddings.word_embeddings.weight"
    if has_cross and cfg.do_pretrain:
        ties["cls.predictions.decoder.weight"] = "bert.embeddings.word_embeddings.weight"
        ties["cls_visual.predictions.weight"] = "visual.embeddings.word_embeddings.weight"
    return ties


def state_dict_spec(cfg):
    """Ordered [(key, shape)] of the owning (non-alias) tensors for the mode described by `cfg`."""
    spec = [("bert.embeddings.word_embeddings.weight", (VOCAB, H)),
            ("bert.embeddings.position_embeddings.weight", (512, H)),
            ("bert.embeddings.token_type_embeddings.weight", (2, H)),
            ("bert.embeddings.LayerNorm.weight", (H,)), ("bert.embeddings.LayerNorm.bias", (H,))]
    for n in range(cfg.text_num_hidden_layers):
        spec += _encoder_layer_spec("bert.encoder.layer.%d." % n)
    spec += [("bert.pooler.dense.weight", (H, H)), ("bert.pooler.dense.bias", (H,))]
    spec += [("visual.embeddings.word_embeddings.weight", (H, VIDEO_DIM)),
             ("visual.embeddings.word_embeddings.bias", (H,)),
             ("visual.embeddings.position_embeddings.weight", (512, H)),
             ("visual.embeddings.LayerNorm.weight", (H,)), ("visual.embeddings.LayerNorm.bias", (H,))]
    for n in range(cfg.visual_num_hidden_layers):
        spec += _encoder_layer_spec("visual.encoder.layer.%d." % n)
    spec += [("visual.pooler.dense.weight", (H, H)), ("visual.pooler.dense.bias", (H,))]
    has_cross = cfg.stage_two or cfg.train_sim_after_cross
    has_decoder = has_cross and not cfg.train_sim_after_cross
    if has_cross:
        spec += [("cross.embeddings.position_embeddings.weight", (1024, H)),
                 ("cross.embeddings.token_type_embeddings.weight", (2, H)),
                 ("cross.embeddings.LayerNorm.weight", (H,)), ("cross.embeddings.LayerNorm.bias", (H,))]
        for n in range(cfg.cross_num_hidden_layers):
            spec += _encoder_layer_spec("cross.encoder.layer.%d." % n)
        spec += [("cross.pooler.dense.weight", (H, H)), ("cross.pooler.dense.bias", (H,))]
    if has_decoder:
        spec += [("decoder.embeddings.LayerNorm.weight", (H,)), ("decoder.embeddings.LayerNorm.bias", (H,))]
        for n in range(cfg.decoder_num_hidden_layers):
            spec += _decoder_layer_spec("decoder.decoder.layer.%d." % n)
        spec += _mlm_head_spec("decoder.classifier.cls.", VOCAB)
    if has_cross and cfg.do_pretrain:
        spec += _mlm_head_spec("cls.", VOCAB)
        spec += _mlm_head_spec("cls_visual.", VIDEO_DIM)
    if has_cross:
        spec += [("similarity_dense.weight", (1, H)), ("similarity_dense.bias", (1,))]
    spec += [("normalize_video.visual_norm2d.weight", (VIDEO_DIM,)),
             ("normalize_video.visual_norm2d.bias", (VIDEO_DIM,))]
    return spec


def make_state_dict(cfg, seed=0, weight_std=0.04, dtype=torch.float32, init_law=False):
    """Deterministic non-trivial weights: N(0, weight_std) matrices, LayerNorm gains 1 + 0.1 N(0,1), biases 0.02 N(0,1).
    `init_law=True` instead follows the reference's own initialisation (modules/until_module.py:70-85): N(0, 0.02)
    matrices / embedding tables, zero biases, unit LayerNorm gains — BASELINE.json configs[0] "random-init bert-base".

    Each tensor draws from its own generator seeded by (seed, index) so the values do not depend on which other
    tensors a mode contains.  Aliases of tied tensors are added as extra keys sharing storage.
    """
    sd = {}
    for idx, (key, shape) in enumerate(state_dict_spec(cfg)):
        g = torch.Generator().manual_seed(seed * 100003 + _stable_hash(key))
        if key.endswith("LayerNorm.weight") or key.endswith("visual_norm2d.weight"):
            t = torch.ones(shape) if init_law else 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            t = torch.zeros(shape) if init_law else 0.02 * torch.randn(shape, generator=g)
        else:
            t = (0.02 if init_law else weight_std) * torch.randn(shape, generator=g)
        sd[key] = t.to(dtype)
    for alias, owner in tied_keys(cfg).items():
        sd[alias] = sd[owner]
    return sd


def _stable_hash(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def make_batch(cfg, seed=1234, b=None, ragged=True, video_dtype=torch.float32):
    """Synthetic (token-id, 1024-d S3D feature) batch of the shape the reference dataloaders emit
    (dataloaders/dataloader_youcook_retrieval.py:188-189, dataloader_howto100m.py:367-369).

    Returns a dict of CPU tensors keyed by UniVL.forward's argument names (reference modules/modeling.py:188-190).
    """
    g = torch.Generator().manual_seed(seed)
    b = b if b is not None else cfg.batch_size // cfg.n_gpu
    P, W, F = cfg.n_pair, cfg.max_words, cfg.max_frames
    n = b * P

    def randint(lo, hi, shape):
        return torch.randint(lo, hi, shape, generator=g)

    lt = randint(8, W + 1, (n,)) if ragged else torch.full((n,), W)
    lv = randint(4, F + 1, (n,)) if ragged else torch.full((n,), F)
    pos_w = torch.arange(W).unsqueeze(0)
    pos_f = torch.arange(F).unsqueeze(0)
    attention_mask = (pos_w < lt.unsqueeze(1)).long()
    video_mask = (pos_f < lv.unsqueeze(1)).long()
    input_ids = randint(1000, VOCAB, (n, W)) * attention_mask
    input_ids[:, 0] = 101
    input_ids[torch.arange(n), lt - 1] = 102
    video = torch.randn(n, F, VIDEO_DIM, generator=g) * video_mask.unsqueeze(-1)
    batch = dict(input_ids=input_ids.view(b, P, W), token_type_ids=torch.zeros(b, P, W, dtype=torch.long),
                 attention_mask=attention_mask.view(b, P, W), video=video.view(b, P, F, VIDEO_DIM).to(video_dtype),
                 video_mask=video_mask.view(b, P, F))
    if cfg.do_pretrain and cfg.stage_two:
        inner = attention_mask.clone()
        inner[:, 0] = 0
        inner[torch.arange(n), lt - 1] = 0
        pick = (torch.rand(n, W, generator=g) < 0.15) & inner.bool()
        pick[0, 1] = True  # force >= 1 masked token (lt >= 8 so position 1 is an inner token)
        labels = torch.where(pick, input_ids, torch.full_like(input_ids, -1))
        r = torch.rand(n, W, generator=g)
        masked = input_ids.clone()
        masked[pick & (r < 0.8)] = 103
        rnd = randint(1000, VOCAB, (n, W))
        sel = pick & (r >= 0.8) & (r < 0.9)
        masked[sel] = rnd[sel]
        vpick = (torch.rand(n, F, generator=g) < 0.15) & video_mask.bool()
        vpick[0, 0] = True  # force >= 1 masked frame (lv >= 4)
        masked_video = video.clone()
        masked_video[vpick] = 0.0
        vlabels = torch.where(vpick, pos_f.expand(n, F), torch.full((n, F), -1))
        batch.update(pairs_masked_text=masked.view(b, P, W), pairs_token_labels=labels.view(b, P, W),
                     masked_video=masked_video.view(b, P, F, VIDEO_DIM).to(video_dtype),
                     video_labels_index=vlabels.view(b, P, F))
    if cfg.stage_two and (cfg.do_pretrain or cfg.task_type == "caption"):
        lc = randint(4, W, (n,)) if ragged else torch.full((n,), W - 1)
        cap = randint(1000, VOCAB, (n, W))
        cmask = (pos_w < (lc + 1).unsqueeze(1)).long()  # [CLS] + lc tokens as input; lc tokens + [SEP] as output
        inp = torch.zeros(n, W, dtype=torch.long)
        out = torch.zeros(n, W, dtype=torch.long)
        inp[:, 0] = 101
        inp[:, 1:] = cap[:, :-1]
        out[:, :] = cap
        out[torch.arange(n), lc] = 102
        inp = inp * cmask
        out = out * cmask
        # the reference dataloaders pad output ids with 0 and rely on ignore_index=-1 only for MLM; the caption CE
        # therefore also scores pad positions against id 0 (modules/modeling.py:253) — kept as is.
        batch.update(input_caption_ids=inp.view(b, P, W), decoder_mask=cmask.view(b, P, W),
                     output_caption_ids=out.view(b, P, W))
    return batch


# ---- optimizer parity case (tests/golden/ref_bert_adam.pt; generated by oracle/make_golden.py from the reference's
# own BertAdam class) -------------------------------------------------------------------------------------------------
ADAM_CASE_HYPER = (1e-2, 0.1, 0.1, 20)   # lr, coef_lr, warmup, t_total


def adam_case(steps=4):
    """-> (names_shapes, init{name: tensor}, grads[step]{name: tensor}, names that never receive a gradient).
    Names follow the checkpoint layout so the drivers' four parameter groups (main_task_retrieval.py:173-190) form."""
    g = torch.Generator().manual_seed(31)
    names_shapes = [("bert.encoder.layer.0.attention.self.query.weight", (96, 64)),
                    ("bert.encoder.layer.0.attention.self.query.bias", (96,)),
                    ("bert.encoder.layer.0.attention.output.LayerNorm.weight", (64,)),
                    ("bert.pooler.dense.weight", (64, 64)),          # never receives a gradient (grad None)
                    ("visual.encoder.layer.0.intermediate.dense.weight", (200, 64)),
                    ("visual.encoder.layer.0.intermediate.dense.bias", (200,)),
                    ("cross.encoder.layer.1.output.LayerNorm.bias", (64,)),
                    ("similarity_dense.weight", (1, 64)),
                    ("decoder.classifier.cls.predictions.bias", (70001,))]  # > one 64K chunk, odd length
    init = {n: 0.5 * torch.randn(s, generator=g) for n, s in names_shapes}
    no_grad = {"bert.pooler.dense.weight"}
    scales = [3.0, 0.05, 0.4]     # step 1 trips the global clip, step 2 nothing, step 3 the per-tensor clip
    grads = [{n: scales[t % 3] * torch.randn(s, generator=g) * (6.0 if (t == 2 and n.startswith("visual")) else 1.0)
              for n, s in names_shapes if n not in no_grad} for t in range(steps)]
    return names_shapes, init, grads, no_grad
