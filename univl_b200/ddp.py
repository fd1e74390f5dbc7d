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

    def pack(self):
        """fp32 gradients -> bf16 payload (no-op without compression / single process)"""
        if self.compress is None or self.world == 1:
            return
        if self.payload is None:
            self.payload = torch.empty(self.g.numel(), dtype=torch.bfloat16, device=self.g.device)
        if self.g.is_cuda:
            from .runtime import call
            call("univl_cast_f32_to_bf16", self.g.data_ptr(), self.payload.data_ptr(), self.g.numel())
        else:
            self.payload.copy_(self.g)   # gloo logic tests on CPU tensors

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
