"""Build the CUDA model (univl_b200.modules.modeling.UniVL) with synthetic weights for the parity tests."""
import json
import os
import tempfile

import torch

from oracle import synth

BERT_BASE = dict(attention_probs_dropout_prob=0.1, hidden_act="gelu", hidden_dropout_prob=0.1, hidden_size=768,
                 initializer_range=0.02, intermediate_size=3072, max_position_embeddings=512,
                 num_attention_heads=12, num_hidden_layers=12, type_vocab_size=2, vocab_size=30522)
_dir = [None]


def bert_dir():
    if _dir[0] is None:
        d = tempfile.mkdtemp(prefix="univl_bert_base_")
        with open(os.path.join(d, "bert_config.json"), "w") as fh:
            json.dump(BERT_BASE, fh)
        _dir[0] = d
    return _dir[0]


def build_model(cfg, sd=None, seed=0, device="cuda", dropout=0.0):
    from univl_b200.modules.modeling import UniVL
    sd = sd if sd is not None else synth.make_state_dict(cfg, seed=seed)
    model = UniVL.from_pretrained(bert_dir(), "visual-base", "cross-base", "decoder-base",
                                  state_dict={k: v.clone() for k, v in sd.items()}, task_config=cfg)
    if dropout is not None:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = dropout
    model.to(device)
    model.train()
    return model


def to_device(batch, device="cuda"):
    return {k: v.to(device) for k, v in batch.items()}


def grads_by_name(model):
    out, seen = {}, set()
    for name, p in model.named_parameters():
        if p.grad is None or id(p) in seen:
            continue
        seen.add(id(p))
        out[name] = p.grad.detach().float().cpu()
    return out
