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
    """Backward in two phases so the gradient exchange of phase 1 overlaps the compute of phase 2.

    The reference's DDP overlaps bucketed all-reduces with the rest of backward through autograd hooks
    (main_task_retrieval.py:197-198).  Here the step is replayed from CUDA graphs, which cannot contain the NCCL calls,
    so the backward is cut at one tensor instead: the hidden state entering text-encoder layer `split_layer`.
      phase 1  loss -> cross encoder -> visual encoder (side stream) and text layers 11 .. split_layer
               => every gradient of the cross / visual / similarity parameters and of the upper text layers is final
      phase 2  text layers split_layer-1 .. 0 and the text embeddings
    `phase1_ranges` / `phase2_ranges` are the contiguous runs of the flat gradient buffer each phase completes, so the
    caller can all-reduce the first set while phase 2 runs (bench.py).  Needs the flat layout (`FusedBertAdam(model=)`).
    """

    def __init__(self, model, flat, split_layer):
        self.model, self.flat = model, flat
        self.split_layer = int(split_layer)
        stack = model.bert.encoder
        if not 0 < self.split_layer < len(stack.layer):
            raise ValueError("PhasedBackward: split_layer must be inside the text stack")
        stack.__dict__["_split_at"] = self.split_layer
        late = set()
        for name, p in model.named_parameters():
            if name.startswith("bert.embeddings."):
                late.add(id(p))
            elif name.startswith("bert.encoder.layer."):
                if int(name.split(".")[3]) < self.split_layer:
                    late.add(id(p))
        # tied tables (decoder / MLM heads share the word table) receive gradients in phase 1 too: they complete in phase 2
        self.phase2_ranges = self._runs([p for p in flat.params if id(p) in late])
        self.phase1_ranges = self._runs([p for p in flat.params if id(p) not in late])
        # leaves at the far end of each phase's sub-graph: asking autograd for their gradients makes it run the whole
        # branch (the kernels accumulate into the flat buffer themselves and hand autograd None)
        self._far1 = [model.visual.embeddings.position_embeddings.weight, model.normalize_video.visual_norm2d.weight]
        self._far2 = [model.bert.embeddings.word_embeddings.weight]
        self._g_split = None

    def _runs(self, params):
        segs = sorted((self.flat.by_id[id(p)][0], self.flat.by_id[id(p)][1]) for p in params)
        runs = []
        for off, n in segs:
            end = off + (n + 63) // 64 * 64
            if runs and off <= runs[-1][1]:
                runs[-1][1] = max(runs[-1][1], end)
            else:
                runs.append([off, end])
        total = self.flat.total
        return [(a, min(b, total)) for a, b in runs]

    def phase1(self, loss):
        split = self.model.bert.encoder.__dict__.get("_split_tensor")
        if split is None:
            raise RuntimeError("PhasedBackward: the text encoder did not run in this forward")
        grads = torch.autograd.grad(loss, [split] + self._far1, allow_unused=True)
        self._split, self._g_split = split, grads[0]

    def phase2(self):
        torch.autograd.grad([self._split], self._far2, grad_outputs=[self._g_split], allow_unused=True)
        self._split = self._g_split = None
        self.model.bert.encoder.__dict__.pop("_split_tensor", None)
