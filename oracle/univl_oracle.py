"""CPU oracle for the UniVL hot path — a functional restatement of the reference algorithm.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (univl_b200/) never does, and has no CPU path of its own.

What it is: plain PyTorch CPU ops (fp32 by default, fp64 on request) arranged as pure functions over a
`state_dict`, following the reference line by line — every function cites the reference file:line it restates.
It is NOT a copy of the reference's nn.Module classes: there are no modules, no parameters objects, no
configuration classes; the four copy-pasted encoder stacks of the reference collapse into one `encoder_stack`.
Dropout is omitted (p = 0): the reference's dropout masks come from torch's global RNG and cannot be matched by
any other implementation, so parity is defined at p = 0 (SURVEY.md §4 "determinism recipe").

Parity pin: the reference ships no tests or golden vectors, so the pin is
tests/golden/*.pt — outputs of the UNMODIFIED reference (imported from /root/reference in the authoring
container by oracle/make_golden.py) on oracle/synth.py inputs; tests/test_oracle_golden.py checks this oracle
against them on CPU.
"""
import math

import torch
import torch.nn.functional as F

HEADS = 12

# Optional bf16 emulation: round the output (and the incoming gradient) of every linear / LayerNorm / GELU / attention
# context to bf16, i.e. what ANY implementation that keeps activations in bf16 does to the reference algorithm.
# Used by the parity tests to size the error that bf16 storage alone causes on ill-conditioned gradients.
_EMULATE = [False]


class emulate_bf16:
    def __enter__(self):
        self.prev = _EMULATE[0]
        _EMULATE[0] = True

    def __exit__(self, *exc):
        _EMULATE[0] = self.prev
        return False


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _r(x):
    return _RoundBF16.apply(x) if _EMULATE[0] else x


# ---------------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps=1e-12):
    """TF-style LayerNorm, eps inside the sqrt, biased variance (reference modules/until_module.py:49-53)."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return _r(w * ((x - u) / torch.sqrt(s + eps)) + b)


def gelu(x):
    """erf GELU (reference modules/until_module.py:28-33)."""
    return _r(x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))))


def linear(x, sd, pfx):
    w = sd[pfx + ".weight"]
    if _EMULATE[0]:
        w = _RoundBF16.apply(w)
    return _r(F.linear(x, w, sd.get(pfx + ".bias")))


def additive_mask(mask01, dtype):
    """[B,S] 0/1 -> [B,1,1,S] 0/-10000 (reference modules/module_bert.py:429-437, module_visual.py:408-416)."""
    return (1.0 - mask01.to(dtype)).unsqueeze(1).unsqueeze(2) * -10000.0


def multi_head_attention(q_in, kv_in, add_mask, sd, pfx, names=("query", "key", "value")):
    """softmax(QK^T / sqrt(d) + mask) V with separate q/k/v projections
    (reference modules/module_bert.py:171-197; decoder variant module_decoder.py:220-247)."""
    B, Sq, Hd = q_in.shape
    Sk = kv_in.shape[1]
    d = Hd // HEADS

    def split(x, S):
        return x.view(B, S, HEADS, d).permute(0, 2, 1, 3)

    q = split(linear(q_in, sd, pfx + names[0]), Sq)
    k = split(linear(kv_in, sd, pfx + names[1]), Sk)
    v = split(linear(kv_in, sd, pfx + names[2]), Sk)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)  # scale BEFORE the mask add (:182-184)
    scores = scores + add_mask
    probs = torch.softmax(scores, dim=-1)
    ctx = _r(torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous().view(B, Sq, Hd))
    return ctx


def dense_residual_norm(x, residual, sd, pfx):
    """LayerNorm(dense(x) + residual) (reference modules/module_bert.py:207-211, :246-250)."""
    return layer_norm(linear(x, sd, pfx + "dense") + residual, sd[pfx + "LayerNorm.weight"],
                      sd[pfx + "LayerNorm.bias"])


def encoder_layer(x, add_mask, sd, pfx):
    """One BertLayer / VisualLayer / CrossLayer (reference modules/module_bert.py:260-264)."""
    ctx = multi_head_attention(x, x, add_mask, sd, pfx + "attention.self.")
    att = dense_residual_norm(ctx, x, sd, pfx + "attention.output.")
    inter = gelu(linear(att, sd, pfx + "intermediate.dense"))  # :233-236
    return dense_residual_norm(inter, att, sd, pfx + "output.")


def encoder_stack(x, add_mask, sd, pfx, n_layers):
    """reference modules/module_bert.py:273-281 (all layers are kept there; only the last is ever used)."""
    for n in range(n_layers):
        x = encoder_layer(x, add_mask, sd, "%s%d." % (pfx, n))
    return x


def pooler(x, sd, pfx):
    """tanh(dense(h[:, 0])) (reference modules/module_bert.py:290-296, module_cross.py:281-287)."""
    return torch.tanh(linear(x[:, 0], sd, pfx + "dense"))


# ---------------------------------------------------------------------------------------------------------
# the four sub-models
# ---------------------------------------------------------------------------------------------------------
def normalize_video(video, sd):
    """cast -> fp32, view [B,F,1024], LayerNorm(1024) (reference modules/modeling.py:88-92)."""
    w = sd["normalize_video.visual_norm2d.weight"]
    video = torch.as_tensor(video).to(w.dtype)
    video = video.view(-1, video.shape[-2], video.shape[-1])
    return layer_norm(video, w, sd["normalize_video.visual_norm2d.bias"])


def text_encoder(input_ids, token_type_ids, attention_mask, sd, n_layers):
    """BertModel.forward (reference modules/module_bert.py:417-447) with BertEmbeddings (:132-146)."""
    S = input_ids.shape[1]
    pos = torch.arange(S)
    e = (sd["bert.embeddings.word_embeddings.weight"][input_ids]
         + sd["bert.embeddings.position_embeddings.weight"][pos].unsqueeze(0)
         + sd["bert.embeddings.token_type_embeddings.weight"][token_type_ids])
    e = layer_norm(e, sd["bert.embeddings.LayerNorm.weight"], sd["bert.embeddings.LayerNorm.bias"])
    return encoder_stack(e, additive_mask(attention_mask, e.dtype), sd, "bert.encoder.layer.", n_layers)


def visual_encoder(video, video_mask, sd, n_layers):
    """VisualModel.forward (reference modules/module_visual.py:397-425) with VisualEmbeddings (:118-131)."""
    S = video.shape[1]
    e = linear(video, sd, "visual.embeddings.word_embeddings") + \
        sd["visual.embeddings.position_embeddings.weight"][torch.arange(S)].unsqueeze(0)
    e = layer_norm(e, sd["visual.embeddings.LayerNorm.weight"], sd["visual.embeddings.LayerNorm.bias"])
    return encoder_stack(e, additive_mask(video_mask, e.dtype), sd, "visual.encoder.layer.", n_layers)


def cross_encoder(seq_out, vis_out, attention_mask, video_mask, sd, n_layers):
    """UniVL._get_cross_output (reference modules/modeling.py:315-325) + CrossModel.forward
    (modules/module_cross.py:364-394) + CrossEmbeddings (:123-138).  Returns (hidden, pooled, concat_mask)."""
    x = torch.cat((seq_out, vis_out), dim=1)
    concat_mask = torch.cat((attention_mask, video_mask), dim=1)
    types = torch.cat((torch.zeros_like(attention_mask), torch.ones_like(video_mask)), dim=1)
    S = x.shape[1]
    e = x + sd["cross.embeddings.position_embeddings.weight"][torch.arange(S)].unsqueeze(0) + \
        sd["cross.embeddings.token_type_embeddings.weight"][types]
    e = layer_norm(e, sd["cross.embeddings.LayerNorm.weight"], sd["cross.embeddings.LayerNorm.bias"])
    h = encoder_stack(e, additive_mask(concat_mask, e.dtype), sd, "cross.encoder.layer.", n_layers)
    return h, pooler(h, sd, "cross.pooler."), concat_mask


def lm_head(x, sd, pfx, decoder_weight, transposed=False):
    """LN(gelu(dense(x))) -> tied projection + bias (reference modules/module_bert.py:314-330;
    visual variant multiplies by the UN-transposed [768,1024] input weight, module_visual.py:308-311)."""
    t = gelu(linear(x, sd, pfx + "predictions.transform.dense"))
    t = layer_norm(t, sd[pfx + "predictions.transform.LayerNorm.weight"],
                   sd[pfx + "predictions.transform.LayerNorm.bias"])
    proj = t.matmul(decoder_weight) if transposed else F.linear(t, decoder_weight)
    return proj + sd[pfx + "predictions.bias"]


def caption_decoder(input_caption_ids, enc_out, decoder_mask, enc_mask, sd, n_layers):
    """DecoderModel.forward (reference modules/module_decoder.py:372-406): embeddings (:309-320), causal∧padding
    self-attention mask built as gt(0) * -10000 (:389-396), encoder attention mask (:385-387), layers (:287-292),
    classifier (:342-349).  Returns logits [B, L, vocab]."""
    B, L = input_caption_ids.shape
    dt = enc_out.dtype
    e = sd["bert.embeddings.word_embeddings.weight"][input_caption_ids] + \
        sd["bert.embeddings.position_embeddings.weight"][torch.arange(L)].unsqueeze(0)
    x = layer_norm(e, sd["decoder.embeddings.LayerNorm.weight"], sd["decoder.embeddings.LayerNorm.bias"])
    enc_add = additive_mask(enc_mask, dt)
    answer = decoder_mask.to(dt).unsqueeze(1).unsqueeze(2)
    future = torch.triu(torch.ones(L, L, dtype=dt), diagonal=1).unsqueeze(0).unsqueeze(1)
    slf_add = ((1.0 - answer) + future).gt(0).to(dt) * -10000.0
    for n in range(n_layers):
        p = "decoder.decoder.layer.%d." % n
        ctx = multi_head_attention(x, x, slf_add, sd, p + "slf_attn.att.")
        s = dense_residual_norm(ctx, x, sd, p + "slf_attn.output.")
        ctx = multi_head_attention(s, enc_out, enc_add, sd, p + "enc_attn.att.")
        d = dense_residual_norm(ctx, s, sd, p + "enc_attn.output.")
        inter = gelu(linear(d, sd, p + "intermediate.dense"))
        x = dense_residual_norm(inter, d, sd, p + "output.")
    return lm_head(x, sd, "decoder.classifier.cls.", sd["bert.embeddings.word_embeddings.weight"])


# ---------------------------------------------------------------------------------------------------------
# similarity + losses
# ---------------------------------------------------------------------------------------------------------
def mean_pool(seq_out, vis_out, attention_mask, video_mask):
    """reference modules/modeling.py:327-339 (text excludes position 0; guarded video denominator)."""
    am = attention_mask.to(seq_out.dtype).unsqueeze(-1).clone()
    am[:, 0, :] = 0.0
    text = (seq_out * am).sum(1) / am.sum(1)
    vm = video_mask.to(vis_out.dtype).unsqueeze(-1)
    den = vm.sum(1)
    den = torch.where(den == 0.0, torch.ones_like(den), den)
    vid = (vis_out * vm).sum(1) / den
    return text, vid


def cross_similarity(seq_out, vis_out, attention_mask, video_mask, sd, n_layers):
    """All (text i, video j) pairs through the cross encoder, pooled -> similarity_dense
    (reference modules/modeling.py:341-375; the 5-row chunking there is only a memory device)."""
    bt, bv = seq_out.shape[0], vis_out.shape[0]
    rows = []
    for i in range(bt):
        s = seq_out[i:i + 1].expand(bv, -1, -1)
        m = attention_mask[i:i + 1].expand(bv, -1)
        _, pooled, _ = cross_encoder(s, vis_out, m, video_mask, sd, n_layers)
        rows.append(linear(pooled, sd, "similarity_dense").squeeze(-1))
    return torch.stack(rows, 0)


def similarity_logits(seq_out, vis_out, attention_mask, video_mask, sd, cfg, pretrain_joint=False):
    """reference modules/modeling.py:377-391."""
    stage_two = bool(getattr(cfg, "stage_two", False))
    align = (not stage_two) and bool(getattr(cfg, "train_sim_after_cross", False))
    if (stage_two and not pretrain_joint) or align:
        return cross_similarity(seq_out, vis_out, attention_mask, video_mask, sd, cfg.cross_num_hidden_layers)
    t, v = mean_pool(seq_out, vis_out, attention_mask, video_mask)
    if not cfg.use_mil:
        t = F.normalize(t, dim=-1)
        v = F.normalize(v, dim=-1)
    return t.matmul(v.t())


def max_margin_loss(sim, cfg):
    """reference modules/until_module.py:223-251 (diagonal included in the mean; weighting only if n_pair>1)."""
    d = torch.diag(sim)
    mm = F.relu(cfg.margin + sim - d.view(-1, 1)) + F.relu(cfg.margin + sim - d.view(1, -1))
    bs = cfg.batch_size // cfg.n_gpu
    if cfg.negative_weighting and cfg.n_pair > 1 and bs > 1:
        easy = 1 - cfg.hard_negative_rate
        alpha = easy / ((bs - 1) * (1 - easy))
        mask = (1 - alpha) * torch.eye(bs, dtype=torch.float64) + alpha
        mask = torch.kron(mask, torch.ones(cfg.n_pair, cfg.n_pair, dtype=torch.float64))
        mask = (mask * (bs * (1 - easy))).to(sim.dtype)
        mm = mm * mask
    return mm.mean()


def cross_en_loss(sim):
    """reference modules/until_module.py:182-191."""
    return -torch.diag(F.log_softmax(sim, dim=-1)).mean()


def mil_nce_loss(sim, cfg):
    """reference modules/until_module.py:193-221."""
    bs = cfg.batch_size // cfg.n_gpu
    P = cfg.n_pair
    mask = torch.kron(torch.eye(bs, dtype=torch.float64), torch.ones(P, P, dtype=torch.float64)).to(sim.dtype)
    from_text = sim + mask * -1e12
    new_sim = torch.cat([sim.t(), from_text], dim=-1)
    logpt = F.log_softmax(new_sim, dim=-1)
    mask2 = torch.cat([mask, torch.zeros_like(mask)], dim=-1)
    new_logpt = -torch.logsumexp(logpt + (1.0 - mask2) * -1e12, dim=-1)
    pick = torch.arange(bs) * P + P // 2
    return new_logpt[pick].mean()


def mlm_loss(seq_cross, token_labels, sd):
    """reference modules/modeling.py:273-276 (CrossEntropyLoss(ignore_index=-1), :165)."""
    logits = lm_head(seq_cross, sd, "cls.", sd["bert.embeddings.word_embeddings.weight"])
    return F.cross_entropy(logits.view(-1, logits.shape[-1]), token_labels.view(-1), ignore_index=-1)


def mfm_loss(vis_cross, video, video_mask, video_labels_index, sd):
    """reference modules/modeling.py:278-297."""
    scores = lm_head(vis_cross, sd, "cls_visual.", sd["visual.embeddings.word_embeddings.weight"], transposed=True)
    s = scores.view(-1, scores.shape[-1])
    vt = video.permute(2, 0, 1).reshape(video.shape[-1], -1)
    logits = s.mm(vt)
    vm = video_mask.to(s.dtype).view(-1)
    masked = logits + (1.0 - vm.view(-1, 1) * vm.view(1, -1)) * -1e8
    nce = -torch.diag(F.log_softmax(masked, dim=-1))
    return nce[(video_labels_index != -1).view(-1)].mean()


# ---------------------------------------------------------------------------------------------------------
# UniVL.forward
# ---------------------------------------------------------------------------------------------------------
def sequence_visual_output(batch_ids, batch_types, batch_mask, video_norm, video_mask, sd, cfg):
    """reference modules/modeling.py:299-313 (inputs already flattened / normalised: `shaped=True`)."""
    seq = text_encoder(batch_ids, batch_types, batch_mask, sd, cfg.text_num_hidden_layers)
    vis = visual_encoder(video_norm, video_mask, sd, cfg.visual_num_hidden_layers)
    return seq, vis


def univl_forward(sd, cfg, batch, return_parts=False):
    """Training-mode UniVL.forward (reference modules/modeling.py:188-271): the stage-dependent sum of up to five
    losses.  `batch` holds the reference's keyword arguments (oracle/synth.py:make_batch)."""
    flat = lambda t: t.view(-1, t.shape[-1])  # noqa: E731
    ids, types, am = flat(batch["input_ids"]), flat(batch["token_type_ids"]), flat(batch["attention_mask"])
    vm = flat(batch["video_mask"])
    video = normalize_video(batch["video"], sd)
    seq, vis = sequence_visual_output(ids, types, am, video, vm, sd, cfg)
    stage_two = bool(getattr(cfg, "stage_two", False))
    parts = {}
    loss = 0.0
    sim_loss_fn = (lambda s: mil_nce_loss(s, cfg)) if cfg.use_mil else (lambda s: max_margin_loss(s, cfg))
    if not stage_two:
        sim = similarity_logits(seq, vis, am, vm, sd, cfg)
        parts["sim_matrix"] = sim
        parts["sim_loss"] = sim_loss_fn(sim)
        loss = loss + parts["sim_loss"]
    else:
        seq_a, vis_a = seq, vis
        if cfg.do_pretrain:
            mids, labels = flat(batch["pairs_masked_text"]), flat(batch["pairs_token_labels"])
            mvideo = normalize_video(batch["masked_video"], sd)
            vlab = flat(batch["video_labels_index"])
            seq_a, vis_a = sequence_visual_output(mids, types, am, mvideo, vm, sd, cfg)
            cross, _, _ = cross_encoder(seq_a, vis_a, am, vm, sd, cfg.cross_num_hidden_layers)
            W = am.shape[-1]
            parts["mlm_loss"] = mlm_loss(cross[:, :W], labels, sd)
            parts["mfm_loss"] = mfm_loss(cross[:, W:], video, vm, vlab, sd)
            joint = similarity_logits(seq, vis, am, vm, sd, cfg, pretrain_joint=True)
            parts["joint_sim_loss"] = sim_loss_fn(joint)
            loss = loss + parts["mlm_loss"] + parts["mfm_loss"] + parts["joint_sim_loss"]
        if batch.get("input_caption_ids") is not None and (cfg.do_pretrain or cfg.task_type == "caption"):
            cin, cmask = flat(batch["input_caption_ids"]), flat(batch["decoder_mask"])
            cout = flat(batch["output_caption_ids"])
            cross, _, cmask_enc = cross_encoder(seq_a, vis_a, am, vm, sd, cfg.cross_num_hidden_layers)
            logits = caption_decoder(cin, cross, cmask, cmask_enc, sd, cfg.decoder_num_hidden_layers)
            parts["decoder_logits"] = logits
            parts["decoder_loss"] = F.cross_entropy(logits.view(-1, logits.shape[-1]), cout.view(-1),
                                                    ignore_index=-1)
            loss = loss + parts["decoder_loss"]
        if cfg.do_pretrain or cfg.task_type == "retrieval":
            sim = similarity_logits(seq_a, vis_a, am, vm, sd, cfg)
            parts["sim_matrix"] = sim
            parts["cross_sim_loss"] = cross_en_loss(sim)
            loss = loss + parts["cross_sim_loss"]
    parts["sequence_output"] = seq
    parts["visual_output"] = vis
    parts["loss"] = loss
    return (loss, parts) if return_parts else loss
