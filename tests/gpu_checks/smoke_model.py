"""GPU bring-up: run golden cases once and print diagnostics (not a pytest file)."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import synth  # noqa: E402
from tests.model_util import build_model, grads_by_name, to_device  # noqa: E402
from tests.oracle_util import load_golden, run_oracle  # noqa: E402

names = sys.argv[1:] or ["ft_joint_npair2", "pretrain1_mil", "caption_small", "pretrain2_small"]
for name in names:
    try:
        gold = load_golden(name)
        cfg = synth.task_config(**gold["cfg_kwargs"])
        sd = synth.make_state_dict(cfg, seed=gold["weight_seed"], **gold.get("weight_kwargs", {}))
        batch = synth.make_batch(cfg, **gold["batch_kwargs"])
        model = build_model(cfg, sd=sd)
        loss = model(**to_device(batch))
        loss.backward()
        torch.cuda.synchronize()
        got = float(loss.detach())
        print("CASE", name, "loss", got, "ref", gold["loss"], "diff", got - gold["loss"], flush=True)
        o_loss, parts, o_grads = run_oracle(cfg, batch, sd=sd, backward=True)
        grads = grads_by_name(model)
        biggest = max(gold["grad_norms"].values())
        rows = []
        num = den = gg = 0.0
        for k, rn in gold["grad_norms"].items():
            g, r = grads[k].double(), o_grads[k].double()
            err = float((g - r).norm())
            rows.append((err / max(rn, 0.05 * biggest), err, rn, float(g.norm()), k))
            num += float((g.flatten() @ r.flatten()))
            den += float(r.norm()) ** 2
            gg += float(g.norm()) ** 2
        print("   global cosine %.5f   biggest ref norm %.3e" % (num / (den ** 0.5 * gg ** 0.5), biggest))
        rows.sort(reverse=True)
        for rel, err, rn, n, k in rows[:12]:
            print("   rel %.3f  err %.3e  ref %.3e  ours %.3e  %s" % (rel, err, rn, n, k))
        if "sim_matrix" in parts:
            print("   oracle sim:", parts["sim_matrix"].detach().flatten()[:8].tolist())
    except Exception:
        traceback.print_exc()
        print("CASE", name, "EXCEPTION", flush=True)
        break
