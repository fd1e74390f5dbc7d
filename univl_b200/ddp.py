"""Data-parallel gradient exchange for the flat gradient buffer.

The reference's only collective is DistributedDataParallel's bucketed gradient all-reduce (main_task_retrieval.py:
197-198).  Here gradients already live in ONE flat fp32 buffer (univl_b200/optim.py), so the exchange is an in-place
NCCL all-reduce over NVLink 5 / NVSwitch of `n_buckets` contiguous slices — no flatten/unflatten copies, no per-
parameter hooks, no unused-parameter graph walk (unused parameters simply contribute zeros).  The mean (1/world) is
folded into the optimizer's gradient read (`FusedBertAdam(grad_scale=1/world)`).  Parameters are broadcast from rank
0 once at construction, as DDP does.  Works on `gloo` for the CPU logic tests.
"""
import torch
import torch.distributed as dist


class FlatGradReducer:
    """compress="bf16": the gradients travel as bf16 (half the NVLink bytes; SURVEY.md §8d/e "bf16 flat buckets") —
    `pack()` rounds the fp32 buffer into a bf16 payload (csrc/misc.cu cast kernels; capturable into the backward CUDA
    graph), `all_reduce()` sums the payload, `unpack()` expands the sum back into the fp32 buffer the optimizer reads."""

    def __init__(self, flat_params, flat_grads, n_buckets=4, group=None, broadcast=True, compress=None):
        self.p, self.g = flat_params, flat_grads
        if compress not in (None, "bf16"):
            raise ValueError("FlatGradReducer: compress must be None or 'bf16'")
        self.compress = compress
        self.payload = None
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = self.g.numel()
        n_buckets = max(1, min(n_buckets, n))
        step = (n + n_buckets - 1) // n_buckets
        step = (step + 1023) // 1024 * 1024
        self.slices = [(i, min(n, i + step)) for i in range(0, n, step)]
        if broadcast and self.world > 1:
            dist.broadcast(self.p, src=0, group=group)

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def pack(self, ranges=None):
        """fp32 gradients -> bf16 payload (no-op without compression / single process); `ranges`: only these [a, b)"""
        if self.compress is None or self.world == 1:
            return
        if self.payload is None:
            self.payload = torch.empty(self.g.numel(), dtype=torch.bfloat16, device=self.g.device)
        for a, b in (ranges if ranges is not None else [(0, self.g.numel())]):
            if self.g.is_cuda:
                from .runtime import call
                call("univl_cast_f32_to_bf16", self.g[a:b].data_ptr(), self.payload[a:b].data_ptr(), b - a)
            else:
                self.payload[a:b].copy_(self.g[a:b])   # gloo logic tests on CPU tensors

    def all_reduce_ranges(self, ranges, async_op=True):
        """sum-reduce the given [a, b) runs of the payload (bf16) or of the gradient buffer; returns the work handles"""
        if self.world == 1:
            return []
        buf = self.payload if self.compress is not None else self.g
        return [dist.all_reduce(buf[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
                for a, b in ranges]

    def unpack(self):
        """summed bf16 payload -> fp32 gradient buffer"""
        if self.compress is None or self.world == 1:
            return
        if self.g.is_cuda:
            from .runtime import call
            call("univl_cast_bf16_to_f32", self.payload.data_ptr(), self.g.data_ptr(), self.g.numel())
        else:
            self.g.copy_(self.payload)

    def all_reduce(self, async_op=False, packed=False):
        """sum-reduce every bucket in place; returns the work handles when async.  With compression, `packed=True`
        means pack() already ran (e.g. inside the captured backward graph) and the caller runs unpack() itself."""
        if self.world == 1:
            return []
        buf = self.g
        if self.compress is not None:
            if not packed:
                self.pack()
            buf = self.payload
        works = []
        for a, b in self.slices:
            w = dist.all_reduce(buf[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if async_op:
                works.append(w)
        if self.compress is not None and not packed and not async_op:
            self.unpack()
        return works


class PhasedBackward:
    """Backward in phases so the gradient exchange of one phase overlaps the compute of the next.

    The reference's DDP overlaps bucketed all-reduces with the rest of backward through autograd hooks
    (main_task_retrieval.py:197-198).  Here the step is replayed from CUDA graphs, which cannot contain the NCCL calls,
    so the autograd graph is cut instead, at the hidden states entering the text-encoder layers `cut_layers`
    (descending, e.g. (9, 5)): the forward continues from a detached leaf there (modules/transformer.py EncoderStack.run).
      phase 0   loss.backward(): heads, cross encoder, decoder, visual encoder (side stream), text layers 11 .. c0
      phase i   text layers c(i-1)-1 .. c(i)          (seeded with the gradient phase i-1 left on its cut leaf)
      last      text layers c(last)-1 .. 0 and the text embeddings (the word table is tied to the decoder / MLM heads,
                so its gradient is only complete here)
    `ranges[i]` are the contiguous runs of the flat gradient buffer that are final once phase i has run; the caller
    all-reduces them while phase i+1 computes (bench.py).  Parameters outside the 2-D weight arena (biases, LayerNorm
    vectors, 0.2 % of the bytes) go with the last phase.  Needs the flat layout (`FusedBertAdam(..., model=model)`).
    """

    def __init__(self, model, flat, cut_layers):
        self.model, self.flat = model, flat
        cuts = sorted({int(c) for c in cut_layers}, reverse=True)
        stack = model.bert.encoder
        if not cuts or not all(0 < c < len(stack.layer) for c in cuts):
            raise ValueError("PhasedBackward: cut layers must lie inside the text stack, got %r" % (cut_layers,))
        self.cuts = cuts
        self.n_phases = len(cuts) + 1
        stack.__dict__["_cut_layers"] = frozenset(cuts)
        arena_end = flat.arena.buf.numel() if flat.arena.entries else 0
        last = self.n_phases - 1
        label = {}
        for name, p in model.named_parameters():
            ph = 0
            if name.startswith("bert.embeddings."):
                ph = last
            elif name.startswith("bert.encoder.layer."):
                li = int(name.split(".")[3])
                ph = sum(1 for c in cuts if li < c)
            if flat.by_id[id(p)][0] >= arena_end:
                ph = last
            label[id(p)] = max(ph, label.get(id(p), 0))          # tied parameters: the latest phase that touches them
        segs = sorted((flat.by_id[id(p)][0], flat.by_id[id(p)][1], label[id(p)]) for p in flat.params)
        self.ranges = [[] for _ in range(self.n_phases)]
        prev = None
        for off, n, ph in segs:
            if ph == prev:                      # neighbour in the buffer, same phase: extend the run (gap included)
                self.ranges[ph][-1][1] = off + n
            else:
                self.ranges[ph].append([off, off + n])
            prev = ph
        # pad each run to its 64-element slot (the padding holds zeros; keeps bf16 payload slices 128-byte aligned)
        self.ranges = [[(a, min((b + 63) // 64 * 64, flat.total)) for a, b in runs] for runs in self.ranges]
        self._pairs = None

    def covered(self):
        """total elements in all ranges (== flat.total when every slot is covered; asserted by the tests)"""
        return sum(b - a for runs in self.ranges for a, b in runs)

    def _stack(self):
        return self.model.bert.encoder.__dict__

    def begin(self):
        """call before the forward of every step (drops the cut tensors of the previous one)"""
        self._stack().pop("_cut_pairs", None)

    def remove(self):
        """take the cuts out of the text stack again: later forwards build one autograd graph, as before"""
        self._stack().pop("_cut_layers", None)
        self._stack().pop("_cut_pairs", None)
        self._pairs = None

    def backward(self, phase, loss=None):
        """run phase `phase` (0 takes the loss)"""
        if phase == 0:
            pairs = self._stack().pop("_cut_pairs", None)
            if not pairs:
                raise RuntimeError("PhasedBackward: the text encoder did not run (or ran without grad) in this forward")
            self._pairs = pairs
            loss.backward()
            return
        cut = self.cuts[phase - 1]
        outs = [x for i, x, leaf in self._pairs if i == cut]
        grads = [leaf.grad for i, x, leaf in self._pairs if i == cut]
        if any(g is None for g in grads):
            raise RuntimeError("PhasedBackward: no gradient reached the cut at text layer %d" % cut)
        torch.autograd.backward(outs, grads)
        if phase == self.n_phases - 1:
            self._pairs = None
