"""GPU: the CUDA path (through the reference-facing UniVL class, down the C ABI) against
 (a) the golden outputs of the unmodified reference (tests/golden/, made by oracle/make_golden.py) and
 (b) the CPU oracle on the same seeded inputs — loss, hidden states and every gradient tensor.

Stated tolerances (bf16 activations, fp32 accumulation / statistics / losses; the reference is fp32):
  loss   retrieval loss with the reference's own random-init law (BASELINE.json configs[0]; *_init cases):
                                                                  |d| <= 1e-3                  (north_star)
         retrieval loss on the stress weights (2x init std, random LayerNorm gains/biases): |d| <= 4e-3
         cross-entropy losses over the vocabulary / frames (caption, pretrain-II):            |d| <= 2e-3 * |loss|
         MIL-NCE on UN-normalised dot products (use_mil): |d| <= 2^-8 * max|sim| — the logits are only resolved to
         bf16 precision relative to their magnitude (~25 here)
  hidden states   max abs err <= 6e-2 on values of magnitude ~4
  gradients       per tensor ||g - g_oracle|| <= 0.10 * max(||g_oracle||, 0.05 * largest gradient norm) and, for tensors
                  above that floor, norm within 6 %.  The floor exists because some gradients are mathematically (key
                  biases: softmax shift invariance) or numerically (q/k weights of deep layers once attention has
                  become uniform: 1e-4 of the value-weight gradients) zero; there only the absolute error is
                  meaningful in bf16.  A tensor that misses these bounds is still accepted if its error is within 2.5x of
                  the error the CPU oracle itself makes when its activations / activation-gradients are rounded to
                  bf16 (oracle.univl_oracle.emulate_bf16) — the conditioning of the gradient, not the kernels.
"""
import pytest
import torch

from oracle import synth
from tests.model_util import build_model, grads_by_name, to_device
from tests.oracle_util import load_golden, run_oracle

pytestmark = pytest.mark.gpu

CASES = ["ft_joint_npair2", "pretrain1_mil", "caption_small", "pretrain2_small", "cfg1_ft_joint", "cfg1_ft_align",
         "cfg1_ft_joint_init", "cfg1_ft_align_init"]


def loss_tolerance(cfg, gold):
    init_law = bool(gold.get("weight_kwargs", {}).get("init_law"))
    if cfg.mode in ("ft_joint", "ft_align"):
        return 1e-3 if (init_law or cfg.mode == "ft_joint") else 4e-3
    if cfg.use_mil and not getattr(cfg, "stage_two", False):
        return 2.0 ** -8 * max(float(s.abs().max()) for s in gold["sim_matrices"])
    return 2e-3 * abs(gold["loss"])


@pytest.mark.parametrize("name", CASES)
def test_loss_hidden_and_grads_match_reference(name):
    gold = load_golden(name)
    cfg = synth.task_config(**gold["cfg_kwargs"])
    wkw = gold.get("weight_kwargs", {})
    sd = synth.make_state_dict(cfg, seed=gold["weight_seed"], **wkw)
    batch = synth.make_batch(cfg, **gold["batch_kwargs"])
    model = build_model(cfg, sd=sd)
    loss = model(**to_device(batch))
    loss.backward()
    torch.cuda.synchronize()
    got = float(loss.detach())
    tol = loss_tolerance(cfg, gold)
    assert abs(got - gold["loss"]) <= tol, "loss %r vs reference %r (tol %g)" % (got, gold["loss"], tol)

    o_loss, parts, o_grads = run_oracle(cfg, batch, sd=sd, backward=True)
    assert abs(got - float(o_loss)) <= tol
    model.eval()
    with torch.no_grad():
        b = to_device(batch)
        seq, vis = model.get_sequence_visual_output(b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                                    b["video"], b["video_mask"])
    seq, vis = seq.float().cpu(), vis.float().cpu()
    # hidden states after a 12-layer bf16 stack against the fp32 reference: values are O(1) (LayerNorm outputs), the
    # bf16 storage rounding alone is 2^-9 relative per layer.  Bound the bulk (relative Frobenius error) tightly and the
    # worst single element (a max over 10^4-10^5 elements, i.e. the noise tail) at ~6 bf16 ulps of |x| = 2..4.
    for ours, ref in ((seq, parts["sequence_output"].detach()), (vis, parts["visual_output"].detach())):
        assert float((ours - ref).norm() / ref.norm()) <= 2e-2  # 1.2e-2 measured on the 2x-init-std stress weights
        assert float((ours - ref).abs().max()) <= 1e-1
    assert (seq[:, :6, :16] - gold["seq_slice"]).abs().max() <= 1e-1

    grads = grads_by_name(model)
    assert set(grads) == set(gold["grad_norms"]), sorted(set(grads) ^ set(gold["grad_norms"]))[:8]
    biggest = max(gold["grad_norms"].values())
    floor = 0.05 * biggest
    bad = []
    emu = None
    for k, ref_norm in gold["grad_norms"].items():
        g, r = grads[k].double(), o_grads[k].double()
        abs_err = float((g - r).norm())
        err = abs_err / max(float(r.norm()), floor)
        ratio = float(g.norm()) / max(ref_norm, 1e-30)
        if err <= 0.10 and (ref_norm < floor or 0.94 <= ratio <= 1.06):
            continue
        # ill-conditioned gradient (e.g. the all-pairs hinge loss at random init: d loss / d sim sums to zero while
        # every pair back-propagates nearly the same vector).  Budget: what rounding the reference algorithm's own
        # activations and activation-gradients to bf16 does (oracle bf16 emulation), with a 2.5x allowance.
        if emu is None:
            _, _, emu = run_oracle(cfg, batch, sd=sd, backward=True, bf16_emulation=True)
        emu_err = float((emu[k].double() - r).norm())
        if abs_err > 2.5 * emu_err:
            bad.append((k, round(err, 4), round(ratio, 4), "bf16-emulated oracle error %.3e vs ours %.3e" % (
                emu_err, abs_err)))
    assert not bad, "gradient mismatches (name, relative error, norm ratio): %s" % bad[:12]
