"""CPU-only checks: the C-ABI library loads without a GPU and exports every symbol include/univl_b200.h declares,
the reference-surface modules reproduce the checkpoint layout contract, and the product refuses to run without CUDA
(no CPU fallback)."""
import ctypes
import os
import sys

import pytest
import torch

from oracle import synth
from tests.model_util import build_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    from univl_b200 import build, lib
    path = build.build()
    assert os.path.exists(path)
    declared = lib.parse_header()
    assert len(declared) >= 30
    handle = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(handle, name), "missing export: " + name
    h = lib.load()
    assert h.univl_abi_version() == 1
    assert h.univl_last_error_string() is not None


def test_argument_errors_are_reported_not_swallowed():
    """bad arguments fail before any launch: usable without a GPU"""
    from univl_b200 import lib
    with pytest.raises(RuntimeError, match="gemm"):
        lib.call("univl_gemm_bf16", None, 0, 0, None, 0, 0, 0, 0, 0, None, 0, 0, None, None, 0, None, 0, 1.0, 0, 0,
                 None)
    with pytest.raises(RuntimeError, match="multiple of 256"):
        lib.call("univl_layernorm_fwd", 1, None, 1, 1, 1, None, None, 4, 100, 1e-12, 0.0, 0, 0, 0, None)
    with pytest.raises(RuntimeError, match="S <= 256"):
        lib.call("univl_attention_fwd", 16, 768, 16, 768, 16, 768, 16, 768, None, None, None, 0, 0, 0, 0, 1, 12, 300,
                 300, 0, 0.125, 0.0, 0, 0, None)


@pytest.mark.parametrize("mode,n_keys", [("ft_joint", 304), ("ft_align", 344), ("caption", 432), ("pretrain2", 444)])
def test_state_dict_layout_contract(mode, n_keys):
    """key names, shapes and tied storage of SURVEY.md Appendix A (full-depth 12/6/2/3 models)."""
    cfg = synth.task_config(mode=mode)
    spec = dict(synth.state_dict_spec(cfg))
    ties = synth.tied_keys(cfg)
    from univl_b200.modules.modeling import UniVL
    from tests.model_util import bert_dir
    model = UniVL.from_pretrained(bert_dir(), "visual-base", "cross-base", "decoder-base", task_config=cfg)
    sd = model.state_dict()
    assert len(sd) == n_keys
    assert set(sd) == set(spec) | set(ties)
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
        assert sd[k].dtype == torch.float32
    for alias, owner in ties.items():
        assert sd[alias].data_ptr() == sd[owner].data_ptr(), alias
    # reference init law: N(0, 0.02) matrices, zero biases, unit LayerNorm gains (until_module.py:70-85)
    w = sd["bert.encoder.layer.3.intermediate.dense.weight"]
    assert abs(float(w.std()) - 0.02) < 1e-3 and abs(float(w.mean())) < 1e-3
    assert float(sd["bert.encoder.layer.3.intermediate.dense.bias"].abs().max()) == 0.0
    assert torch.equal(sd["visual.embeddings.LayerNorm.weight"], torch.ones(768))
    # optimizer grouping of the reference driver keys on these substrings (main_task_retrieval.py:173-190)
    names = [n for n, _ in model.named_parameters()]
    assert any("bert." in n and "LayerNorm.weight" in n for n in names)
    unused = {"bert.pooler.dense.weight", "bert.pooler.dense.bias", "visual.pooler.dense.weight",
              "visual.pooler.dense.bias"}
    assert unused <= set(names)


def test_checkpoint_round_trip_and_gamma_beta_rename():
    cfg = synth.task_config(mode="ft_joint", text_layers=1, visual_layers=1)
    sd = synth.make_state_dict(cfg)
    old = {}
    for k, v in sd.items():  # TF-style names the loader must rename (until_module.py:94-104)
        old[k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta")] = v
    model = build_model(cfg, sd=old, device="cpu")
    out = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k], v), k
    # unknown / missing keys never raise (non-strict loader)
    extra = dict(sd)
    extra["not.a.real.key"] = torch.zeros(3)
    del extra["bert.pooler.dense.bias"]
    build_model(cfg, sd=extra, device="cpu")


def test_no_cpu_fallback():
    cfg = synth.task_config(mode="ft_joint", text_layers=1, visual_layers=1, batch_size=2)
    model = build_model(cfg, device="cpu")
    batch = synth.make_batch(cfg)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(**batch)
    from univl_b200.modules.until_module import CrossEn, LayerNorm
    with pytest.raises(RuntimeError, match="CUDA"):
        CrossEn()(torch.zeros(3, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        LayerNorm(768)(torch.zeros(2, 768))


def test_config_surface():
    from univl_b200.modules.module_bert import BertConfig
    from univl_b200.modules.module_cross import CrossConfig
    from univl_b200.modules.module_decoder import DecoderConfig
    from univl_b200.modules.module_visual import VisualConfig
    c, sd = VisualConfig.get_config("visual-base", None, 2, None)
    assert c.vocab_size == 1024 and c.num_hidden_layers == 1 and sd is None
    c, _ = CrossConfig.get_config("cross-base", None, 2, None)
    assert c.max_position_embeddings == 1024 and c.type_vocab_size == 2
    c, _ = DecoderConfig.get_config("decoder-base", None, 2, None)
    assert c.max_target_embeddings == 512 and c.num_decoder_layers == 1
    assert BertConfig.get_config("/nonexistent/model", None, 2, None) is None
    b = BertConfig(30522)
    assert b.hidden_size == 768 and "vocab_size" in b.to_json_string()
    with pytest.raises(ValueError):
        BertConfig(3.5)
    from univl_b200.modules.until_module import PreTrainedModel
    with pytest.raises(ValueError):
        PreTrainedModel(object())


def test_launcher_shims(monkeypatch):
    """univl_b200.launcher: --local-rank mapping, import stubs, numpy aliases, `modules` shadowing"""
    import importlib
    import numpy as np
    from univl_b200 import launcher
    assert launcher.fix_rank_args(["--do_train", "--local-rank=3"]) == ["--do_train", "--local_rank=3"]
    monkeypatch.setenv("LOCAL_RANK", "5")
    assert launcher.fix_rank_args(["--do_train"]) == ["--do_train", "--local_rank", "5"]
    assert launcher.fix_rank_args(["--local_rank", "1"]) == ["--local_rank", "1"]
    launcher.install_stubs()
    import boto3  # noqa: F401
    from botocore.exceptions import ClientError  # noqa: F401
    import nlgeval
    assert hasattr(nlgeval, "NLGEval")
    launcher.install_numpy_aliases()
    assert np.zeros(2, dtype=np.float).dtype == np.float64
    from oracle import build_ref
    root = build_ref.ref_root()
    if root is None:
        return
    saved = {k: v for k, v in sys.modules.items() if k == "modules" or k.startswith("modules.")}
    saved_path = list(sys.path)
    try:
        launcher.install_shadow(root)
        m = importlib.import_module("modules.modeling")
        assert m.__name__ == "univl_b200.modules.modeling"
        from modules.optimization import BertAdam
        from univl_b200.optim import FusedBertAdam
        assert issubclass(BertAdam, FusedBertAdam)
        tok = importlib.import_module("modules.tokenization")          # the checkout's own file, our file_utils under it
        assert os.path.dirname(tok.__file__) == os.path.join(root, "modules")
        assert hasattr(tok, "BertTokenizer")
        from modules.beam import Beam  # noqa: F401
    finally:
        for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = saved_path
