"""KV-cached caption decoding and beam search.

The reference's caption evaluation (main_task_caption.py:434-517) calls `model.decoder_caption` once per generated
token on the FULL prefix: every step re-runs the 2-layer cross encoder on n_inst * n_beam copies of the (text, video)
pair and the 3 decoder layers on all t prefix tokens, then keeps only the last position's logits — O(L^2) decoder work
and L * n_beam redundant cross-encoder passes.  Here, per batch of instances:

  * the cross encoder runs ONCE per instance; the key/value projections of every decoder layer's encoder-attention are
    computed ONCE per instance and shared by its beams — the n_beam hypotheses of an instance are the Sq = n_beam query
    rows of one attention "sequence" over the instance's S_e encoder keys, so nothing is ever `repeat`ed;
  * each decoder layer keeps a self-attention K/V cache [n_inst * n_beam, L_max, 2H]; a step projects only the NEW
    token (q | k | v in one GEMM), appends its k | v at position t, and attends with Sq = 1 over the cache under a
    (position <= t) mask; beam reordering is an index_select of the cache rows by the beams' back-pointers;
  * the vocabulary projection runs on the n_inst * n_beam last-token rows only.

Same kernels as the training path (C ABI: univl_gemm_bf16, univl_attention_fwd, univl_layernorm_fwd,
univl_embed_text_fwd); arithmetic per token is identical to `decoder_caption` on the full prefix (tested to bf16
tolerance against it).  The beam bookkeeping restates `modules/beam.py` (Beam.advance :63-87: scores summed with the
previous beam scores except at the first step, top-k over beam x vocabulary, back-pointers, finished when the top
hypothesis ends in [SEP]) in batched tensor form.
"""
import torch

from . import ops
from . import runtime as rt

BF16 = torch.bfloat16


class CachedCaptionDecoder:
    """Incremental decoder state for `n_inst` instances x `n_beam` hypotheses."""

    def __init__(self, model, sequence_output, visual_output, input_mask, video_mask, n_beam, max_len):
        self.model = model
        dec = model.decoder
        self.dec = dec
        self.n_inst = sequence_output.shape[0]
        self.n_beam = int(n_beam)
        self.max_len = int(max_len)
        self.H = sequence_output.shape[-1]
        dev = sequence_output.device
        self.device = dev
        input_mask = input_mask.reshape(self.n_inst, -1).long().contiguous()
        video_mask = video_mask.reshape(self.n_inst, -1).long().contiguous()
        with rt.use_model(model, dev):
            seq2d = sequence_output.to(BF16).reshape(-1, self.H).contiguous()
            vis2d = visual_output.to(BF16).reshape(-1, self.H).contiguous()
            # reference modeling.py:393-399: decoder attends to the cross-encoder output of (text, video)
            self.enc2d, _, self.Se = model._cross_pairs(seq2d, vis2d, input_mask, video_mask, False)
            arena = rt.current()
            self.enc_mask = ops.MaskSpec(input_mask, video_mask)
            self.enc_kv = []
            for layer in dec.decoder.layer:
                att = layer.enc_attn.att
                wqkv = arena.bf16_qkv(att.query.weight, att.key.weight, att.value.weight)
                self.enc_kv.append(ops.linear_fwd(self.enc2d, wqkv[self.H:], rt.packed_bias(att.key.bias, att.value.bias)))
        rows = self.n_inst * self.n_beam
        self.cache = [torch.zeros(rows, self.max_len, 2 * self.H, dtype=BF16, device=dev) for _ in dec.decoder.layer]
        self.t = 0
        self.active = torch.arange(self.n_inst, device=dev)      # instance index of every active cache slot

    # ------------------------------------------------------------------------------------------------------
    def select(self, keep_positions):
        """keep only these active positions (finished instances leave the batch, main_task_caption.py:400-421)"""
        idx = torch.as_tensor(keep_positions, device=self.device, dtype=torch.long)
        rows = (idx.unsqueeze(1) * self.n_beam + torch.arange(self.n_beam, device=self.device)).reshape(-1)
        self.cache = [c.index_select(0, rows) for c in self.cache]
        self.active = self.active.index_select(0, idx)

    def reorder(self, origin):
        """origin [n_active, n_beam]: beam j of an instance continues hypothesis origin[i, j] of the previous step"""
        n = origin.shape[0]
        rows = (torch.arange(n, device=self.device).unsqueeze(1) * self.n_beam + origin).reshape(-1)
        self.cache = [c.index_select(0, rows) for c in self.cache]

    def step(self, tokens):
        """tokens int64 [n_active * n_beam]: the token at position self.t of every hypothesis.
        -> fp32 logits [n_active * n_beam, vocab] for position self.t + 1"""
        dec, H, t = self.dec, self.H, self.t
        if t >= self.max_len:
            raise RuntimeError("CachedCaptionDecoder: sequence longer than max_len=%d" % self.max_len)
        n_act = self.active.shape[0]
        rows = n_act * self.n_beam
        model = self.model
        with rt.use_model(model, self.device):
            arena = rt.current()
            emb = dec.embeddings
            # word + position[t] -> LayerNorm (reference module_decoder.py:309-320); the position table view starts at row t
            x = ops.EmbedTextFn.apply(tokens.reshape(rows, 1).contiguous(), None, emb.word_embeddings.weight,
                                      emb.position_embeddings.weight[t:], None, emb.LayerNorm.weight,
                                      emb.LayerNorm.bias, 0.0, False)
            pos_mask = (torch.arange(self.max_len, device=self.device) <= t).long().unsqueeze(0).expand(rows, -1)
            slf_mask = ops.MaskSpec(pos_mask.contiguous())
            # encoder K/V and padding masks of the ACTIVE instances (one gather per step; the full set when none left)
            all_active = n_act == self.n_inst
            if all_active:
                enc_mask = self.enc_mask
            else:
                enc_mask = ops.MaskSpec(self.enc_mask.a.index_select(0, self.active),
                                        self.enc_mask.b.index_select(0, self.active))
            for li, layer in enumerate(dec.decoder.layer):
                # ---- causal self-attention of the new token over the cache (module_decoder.py:220-247, :389-396) ----
                att, out = layer.slf_attn.att, layer.slf_attn.output
                wqkv = arena.bf16_qkv(att.query.weight, att.key.weight, att.value.weight)
                qkv = ops.linear_fwd(x, wqkv, rt.packed_bias(att.query.bias, att.key.bias, att.value.bias))
                cache = self.cache[li]
                cache[:, t] = qkv[:, H:]
                kv2d = cache.view(rows * self.max_len, 2 * H)
                ctx, _ = ops.attention_fwd(qkv[:, :H], kv2d[:, :H], kv2d[:, H:], rows, 1, self.max_len, slf_mask)
                ao = ops.linear_fwd(ctx, arena.bf16(out.dense.weight), out.dense.bias)
                x, _, _ = ops.layernorm_fwd(ao, x, out.LayerNorm.weight, out.LayerNorm.bias)
                # ---- encoder attention: the n_beam hypotheses of an instance are n_beam query rows of ONE sequence ----
                att, out = layer.enc_attn.att, layer.enc_attn.output
                wqkv = arena.bf16_qkv(att.query.weight, att.key.weight, att.value.weight)
                q = ops.linear_fwd(x, wqkv[:H], att.query.bias)
                kv = self.enc_kv[li]
                if not all_active:
                    kv = kv.view(self.n_inst, self.Se, 2 * H).index_select(0, self.active).view(-1, 2 * H)
                ctx, _ = ops.attention_fwd(q, kv[:, :H], kv[:, H:], n_act, self.n_beam, self.Se, enc_mask)
                ao = ops.linear_fwd(ctx, arena.bf16(out.dense.weight), out.dense.bias)
                x, _, _ = ops.layernorm_fwd(ao, x, out.LayerNorm.weight, out.LayerNorm.bias)
                # ---- feed-forward ----
                w1, w2 = arena.bf16(layer.intermediate.dense.weight), arena.bf16(layer.output.dense.weight)
                pre = torch.empty(rows, w1.shape[0], dtype=BF16, device=self.device)
                h = ops.linear_fwd(x, w1, layer.intermediate.dense.bias, epi=ops.EPI_GELU, aux_out=pre)
                fo = ops.linear_fwd(h, w2, layer.output.dense.bias)
                x, _, _ = ops.layernorm_fwd(fo, x, layer.output.LayerNorm.weight, layer.output.LayerNorm.bias)
            logits = dec.classifier.cls.logits(x)
        self.t += 1
        return logits


@torch.no_grad()
def beam_search(model, sequence_output, visual_output, input_mask, video_mask, max_words, n_beam=5, bos=101, eos=102):
    """Caption beam search with the semantics of the reference loop (main_task_caption.py:434-517 with modules/beam.py)
    on the KV-cached decoder.  sequence_output [n, W, H], visual_output [n, F, H] (from get_sequence_visual_output).
    -> (hypotheses: list over instances of the best token-id list (without [CLS]), scores: list of floats)"""
    if model.training:
        raise RuntimeError("beam_search: call model.eval() first")
    dev = sequence_output.device
    n_inst = sequence_output.shape[0]
    dec = CachedCaptionDecoder(model, sequence_output, visual_output, input_mask, video_mask, n_beam, max_words)
    scores = torch.zeros(n_inst, n_beam, device=dev)                  # Beam.scores
    next_ys = [torch.full((n_inst, n_beam), bos, dtype=torch.long, device=dev)]
    prev_ks = []
    done = torch.zeros(n_inst, dtype=torch.bool, device=dev)
    active = torch.arange(n_inst, device=dev)                         # instance ids still decoding, in cache order
    for step in range(1, max_words + 1):
        tokens = next_ys[-1].index_select(0, active).reshape(-1)
        logits = dec.step(tokens)
        word_prob = torch.log_softmax(logits, dim=1).view(active.shape[0], n_beam, -1)
        V = word_prob.shape[-1]
        if step == 1:
            beam_lk = word_prob[:, 0]                                 # Beam.advance: only hypothesis 0 exists at first
        else:
            beam_lk = (word_prob + scores.index_select(0, active).unsqueeze(-1)).reshape(active.shape[0], -1)
        best, best_id = beam_lk.topk(n_beam, dim=1, largest=True, sorted=True)
        prev_k = best_id // V
        word = best_id - prev_k * V
        full_k = torch.zeros(n_inst, n_beam, dtype=torch.long, device=dev)
        full_w = torch.zeros(n_inst, n_beam, dtype=torch.long, device=dev)
        full_k[active], full_w[active] = prev_k, word
        scores[active] = best
        prev_ks.append(full_k)
        next_ys.append(full_w)
        finished = word[:, 0] == eos                                  # top-of-beam is [SEP]
        done[active[finished]] = True
        keep = (~finished).nonzero().flatten()
        if keep.numel() == 0:
            break
        dec.reorder(prev_k)
        if keep.numel() != active.numel():
            dec.select(keep)
            active = active.index_select(0, keep)
    # best hypothesis per instance: walk the back-pointers from the top-scoring beam (Beam.sort_scores / get_hypothesis)
    order = scores.argsort(dim=1, descending=True)
    hyps, outs = [], []
    ks = torch.stack(prev_ks).cpu()        # [steps, n_inst, n_beam]
    ys = torch.stack(next_ys).cpu()        # [steps + 1, n_inst, n_beam]
    order_c, scores_c = order.cpu(), scores.cpu()
    n_steps = ks.shape[0]
    # an instance stops advancing at the step it finished: later rows of ks / ys for it are zeros and are not walked
    last = torch.full((n_inst,), n_steps, dtype=torch.long)
    for i in range(n_inst):
        for s in range(n_steps):
            if int(ys[s + 1, i, 0]) == eos:
                last[i] = s + 1
                break
    for i in range(n_inst):
        k = int(order_c[i, 0])
        hyp = []
        for j in range(int(last[i]) - 1, -1, -1):
            hyp.append(int(ys[j + 1, i, k]))
            k = int(ks[j, i, k])
        hyps.append(hyp[::-1])
        outs.append(float(scores_c[i, int(order_c[i, 0])]))
    return hyps, outs
