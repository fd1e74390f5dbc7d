"""GPU: the reference's REAL driver functions (main_task_retrieval.py: init_model :152-167, prep_optimizer :169-200,
train_epoch :318-365, eval_epoch :381-447) running on the univl_b200 package through `univl_b200.launcher`'s shims —
the drop-in claim of INTEGRATION.md §1, exercised rather than asserted.

The unmodified driver is taken from oracle/_ref (staged by oracle/build_ref.py from /root/reference; skipped when
neither exists).  One process, NCCL group of size 1 (the driver calls init_process_group at import, :23), the DDP wrap
with find_unused_parameters=True, clip_grad_norm_, BertAdam.step / zero_grad, get_lr logging, and the eval path.
"""
import argparse
import os
import sys

import pytest
import torch

from oracle import build_ref, synth
from tests.model_util import bert_dir

pytestmark = pytest.mark.gpu


class _Pairs(torch.utils.data.Dataset):
    """items shaped like Youcook_DataLoader.__getitem__ (dataloaders/dataloader_youcook_retrieval.py:188-189)"""

    def __init__(self, cfg, n):
        b = synth.make_batch(cfg, seed=77, b=n)
        self.t = [b["input_ids"], b["attention_mask"], b["token_type_ids"], b["video"].double(), b["video_mask"],
                  b["input_ids"], torch.full_like(b["input_ids"], -1), b["video"].double(),
                  torch.full_like(b["video_mask"], -1)]

    def __len__(self):
        return self.t[0].shape[0]

    def __getitem__(self, i):
        return tuple(t[i] for t in self.t)


@pytest.mark.parametrize("align", [False, True])
def test_reference_retrieval_driver_trains_and_evaluates_on_univl_b200(tmp_path, align):
    root = build_ref.ref_root()
    if root is None:
        pytest.skip("reference checkout not staged (oracle/build_ref.py)")
    from univl_b200 import launcher
    os.environ["MASTER_PORT"] = str(29600 + (os.getpid() % 200))
    launcher.prepare(os.path.join(root, "main_task_retrieval.py"))
    import importlib
    drv = importlib.import_module("main_task_retrieval")      # the unmodified driver; inits NCCL at import
    import modules.modeling
    assert modules.modeling.__name__ == "univl_b200.modules.modeling"
    assert drv.UniVL is modules.modeling.UniVL
    import util
    drv.logger = util.get_logger(str(tmp_path / "log.txt"))

    args = argparse.Namespace(
        do_pretrain=False, do_train=True, do_eval=True, task_type="retrieval", datatype="youcook", stage_two=False,
        train_sim_after_cross=align, batch_size=4, batch_size_val=4, n_gpu=1, n_pair=1, margin=0.1,
        negative_weighting=1, hard_negative_rate=0.5, use_mil=False, sampled_use_mil=False, video_dim=1024,
        max_words=16, max_frames=12, local_rank=0, world_size=1, text_num_hidden_layers=2,
        visual_num_hidden_layers=1, cross_num_hidden_layers=1, decoder_num_hidden_layers=1, init_model=None,
        bert_model=bert_dir(), visual_model="visual-base", cross_model="cross-base", decoder_model="decoder-base",
        cache_dir=str(tmp_path), lr=1e-3, coef_lr=0.1, warmup_proportion=0.1, gradient_accumulation_steps=1,
        n_display=1, epochs=1, output_dir=str(tmp_path), seed=42, fp16=False)
    device = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = drv.init_model(args, device, 1, 0)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    optimizer, scheduler, ddp_model = drv.prep_optimizer(args, model, 12, device, 1, 0, coef_lr=args.coef_lr)
    assert isinstance(ddp_model, torch.nn.parallel.DistributedDataParallel)
    cfg = synth.task_config(mode="ft_align" if align else "ft_joint", batch_size=4, text_layers=2, visual_layers=1,
                            cross_layers=1, max_words=16, max_frames=12)
    loader = torch.utils.data.DataLoader(_Pairs(cfg, 12), batch_size=4, shuffle=False, drop_last=True)
    total, gs = drv.train_epoch(0, args, ddp_model, loader, device, 1, optimizer, scheduler, 0, local_rank=0)
    assert gs == 3 and total == total and total > 0.0                      # 3 optimizer steps, finite mean loss
    moved = [n for n, p in model.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert any("bert.encoder.layer.1" in n for n in moved) and any("visual.encoder" in n for n in moved)
    assert not any("pooler" in n and "cross" not in n for n in moved)       # never receive a gradient
    for p in model.parameters():                                           # zero_grad() cleared what autograd fills
        assert p.grad is None or float(p.grad.abs().max()) == 0.0
    # checkpoint written by the driver loads back through the driver
    out = drv.save_model(0, args, ddp_model, type_name="")
    re = drv.load_model(0, args, 1, device, model_file=out)
    for (n, p), (n2, q) in zip(model.named_parameters(), re.named_parameters()):
        assert n == n2 and torch.equal(p.detach(), q.detach()), n
    # eval path of the driver (all-pairs similarity + retrieval metrics).  With n_gpu == 1 the reference hands
    # compute_metrics a LIST of row blocks (main_task_retrieval.py:443-445: only the n_gpu > 1 branch concatenates) and
    # metrics.py:9 fails on `-x`; that is the reference's own single-GPU bug, so give it the concatenated matrix.
    import numpy as np
    import metrics
    drv.compute_metrics = lambda sm: metrics.compute_metrics(
        np.concatenate(tuple(sm), axis=0) if isinstance(sm, list) else sm)
    r1 = drv.eval_epoch(args, ddp_model, torch.utils.data.DataLoader(_Pairs(cfg, 8), batch_size=4), device, 1)
    assert 0.0 <= float(r1) <= 1.0
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    for name in ("main_task_retrieval",):
        sys.modules.pop(name, None)
