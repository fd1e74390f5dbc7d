"""`BertAdam` with the reference's constructor (modules/optimization.py:66-84) on the fused device kernels.

Drop-in for `from modules.optimization import BertAdam` in the reference drivers: same arguments, same update rule
(see univl_b200/optim.py and csrc/optim.cu for the restated semantics), one fused multi-tensor launch sequence per
step instead of a Python loop over ~300 tensors.  Pass `model=` to also flatten parameters/gradients so the backward
kernels accumulate into the flat gradient buffer and the bf16 weight copies are refreshed by the update kernel.
"""
from ..optim import FusedBertAdam


def warmup_linear(x, warmup=0.002):
    """Triangular schedule of the reference (optimization.py:37-43)."""
    if x < warmup:
        return x / warmup
    return max((x - 1.) / (warmup - 1.), 0)


SCHEDULES = {"warmup_linear": warmup_linear}


class BertAdam(FusedBertAdam):
    def __init__(self, params, lr=None, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 weight_decay=0.01, max_grad_norm=1.0, model=None):
        if lr is None:
            raise ValueError("BertAdam: lr is required")
        super(BertAdam, self).__init__(params, lr=lr, warmup=warmup, t_total=t_total, schedule=schedule, b1=b1,
                                       b2=b2, e=e, weight_decay=weight_decay, max_grad_norm=max_grad_norm,
                                       model=model)

    def get_lr(self):
        """scheduled learning rate per parameter (reference optimization.py:86-101)"""
        if not self._built:
            return [0]
        step = int(self.step_dev.item())
        lrs = []
        for group in self.param_groups:
            for _ in group["params"]:
                if group["t_total"] != -1:
                    lrs.append(group["lr"] * warmup_linear(step / group["t_total"], group["warmup"]))
                else:
                    lrs.append(group["lr"])
        return lrs
