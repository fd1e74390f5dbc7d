"""Configuration objects — same surface as the reference's modules/until_config.py (PretrainedConfig.get_config
:40-99, from_dict :101-107, from_json_file :109-114, to_dict/to_json_string :119-126), re-written.

A config is a bag of attributes loaded from `<dir>/<config_name>`; `get_config` resolves a model name to a directory
next to this file (visual-base, cross-base, decoder-base ship here; bert-base-uncased is supplied by the user exactly
as with the reference) or takes a path as is, and optionally picks up `<dir>/<weights_name>`.
"""
import copy
import json
import logging
import os
import shutil
import tarfile
import tempfile

import torch

from .file_utils import cached_path

logger = logging.getLogger(__name__)


def _log(task_config, level, msg):
    if task_config is None or getattr(task_config, "local_rank", 0) == 0:
        getattr(logger, level)(msg)


class PretrainedConfig(object):
    pretrained_model_archive_map = {}
    config_name = ""
    weights_name = ""

    @classmethod
    def get_config(cls, pretrained_model_name, cache_dir, type_vocab_size, state_dict, task_config=None):
        here = os.path.dirname(os.path.abspath(__file__))
        location = os.path.join(here, pretrained_model_name)  # absolute names pass through unchanged
        if not os.path.exists(location):
            location = cls.pretrained_model_archive_map.get(pretrained_model_name, pretrained_model_name)
        try:
            resolved = cached_path(location, cache_dir=cache_dir)
        except FileNotFoundError:
            _log(task_config, "error", "Model name '{}' was not found; '{}' is not a path or url with a file behind "
                                       "it.".format(pretrained_model_name, location))
            return None
        _log(task_config, "info", "loading archive file {}".format(resolved))
        scratch = None
        if os.path.isdir(resolved):
            folder = resolved
        else:
            scratch = tempfile.mkdtemp()
            with tarfile.open(resolved, "r:gz") as archive:
                archive.extractall(scratch)
            folder = scratch
        try:
            config = cls.from_json_file(os.path.join(folder, cls.config_name))
            config.type_vocab_size = type_vocab_size
            _log(task_config, "info", "Model config {}".format(config))
            if state_dict is None:
                weights = os.path.join(folder, cls.weights_name)
                if os.path.exists(weights):
                    state_dict = torch.load(weights, map_location="cpu")
                else:
                    _log(task_config, "info", "Weight doesn't exsits. {}".format(weights))
        finally:
            if scratch:
                shutil.rmtree(scratch)
        return config, state_dict

    @classmethod
    def from_dict(cls, json_object):
        config = cls(vocab_size_or_config_json_file=-1)
        config.__dict__.update(json_object)
        return config

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding="utf-8") as reader:
            return cls.from_dict(json.load(reader))

    def _init_from(self, first, defaults):
        """shared constructor body: `first` is a json path (str) or the vocab size (int)"""
        if isinstance(first, str):
            with open(first, "r", encoding="utf-8") as reader:
                self.__dict__.update(json.load(reader))
        elif isinstance(first, int):
            self.vocab_size = first
            self.__dict__.update(defaults)
        else:
            raise ValueError("First argument must be either a vocabulary size (int)"
                             "or the path to a pretrained model config file (str)")

    def __repr__(self):
        return str(self.to_json_string())

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"
