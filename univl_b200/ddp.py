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
    def __init__(self, flat_params, flat_grads, n_buckets=4, group=None, broadcast=True):
        self.p, self.g = flat_params, flat_grads
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

    def all_reduce(self, async_op=False):
        """sum-reduce every bucket in place; returns the work handles when async"""
        if self.world == 1:
            return []
        works = []
        for a, b in self.slices:
            w = dist.all_reduce(self.g[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if async_op:
                works.append(w)
        return works
