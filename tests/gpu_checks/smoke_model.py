"""GPU bring-up: run each mode once and print diagnostics (not a pytest file)."""
import sys, os, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import synth
from tests.model_util import build_model, to_device, grads_by_name
from tests.oracle_util import load_golden, run_oracle

names = sys.argv[1:] or ["ft_joint_npair2", "pretrain1_mil", "caption_small", "pretrain2_small"]
for name in names:
    try:
        gold = load_golden(name)
        cfg = synth.task_config(**gold["cfg_kwargs"])
        batch = synth.make_batch(cfg, **gold["batch_kwargs"])
        model = build_model(cfg, seed=gold["weight_seed"])
        loss = model(**to_device(batch))
        loss.backward()
        torch.cuda.synchronize()
        print("CASE", name, "loss", float(loss), "ref", gold["loss"], "diff", float(loss) - gold["loss"], flush=True)
        o_loss, parts, o_grads = run_oracle(cfg, batch, seed=gold["weight_seed"], backward=True)
        model.eval()
        with torch.no_grad():
            b = to_device(batch)
            seq, vis = model.get_sequence_visual_output(b["input_ids"], b["token_type_ids"], b["attention_mask"], b["video"], b["video_mask"])
        print("   seq err", float((seq.float().cpu() - parts["sequence_output"].detach()).abs().max()),
              "vis err", float((vis.float().cpu() - parts["visual_output"].detach()).abs().max()), flush=True)
        grads = grads_by_name(model)
        print("   grad key sets equal:", set(grads) == set(gold["grad_norms"]), sorted(set(grads) ^ set(gold["grad_norms"]))[:6])
        worst = []
        for k, rn in gold["grad_norms"].items():
            if k not in grads: continue
            g = grads[k]; n = float(g.double().norm())
            og = o_grads.get(k)
            cos = float((g.double().flatten() @ og.double().flatten()) / (g.double().norm() * og.double().norm() + 1e-30)) if og is not None else float('nan')
            worst.append((cos, n / (rn + 1e-30), k))
        worst.sort()
        for w in worst[:10]:
            print("   worst cos %.5f ratio %.4f %s" % w)
        print("   median cos %.6f" % worst[len(worst) // 2][0], flush=True)
    except Exception:
        traceback.print_exc()
        print("CASE", name, "EXCEPTION", flush=True)
        break
