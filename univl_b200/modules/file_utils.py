"""Local-only stand-in for the reference's modules/file_utils.py.

The reference module is a download cache (S3/HTTP, imports boto3 at the top — modules/file_utils.py:20-21); it is off
the hot path and out of scope.  The drivers and configs only need these two names:
`PYTORCH_PRETRAINED_BERT_CACHE` (main_task_retrieval.py:17) and `cached_path`, which here resolves local paths and
refuses URLs (there is no network on the target machines).
"""
import os
from pathlib import Path

PYTORCH_PRETRAINED_BERT_CACHE = Path(os.getenv("PYTORCH_PRETRAINED_BERT_CACHE",
                                               Path.home() / ".pytorch_pretrained_bert"))


def cached_path(url_or_filename, cache_dir=None):
    name = str(url_or_filename)
    if name.startswith(("http://", "https://", "s3://")):
        raise EnvironmentError("univl_b200 does not download: fetch {} yourself and pass the local path".format(name))
    if os.path.exists(name):
        return name
    raise FileNotFoundError("file {} not found".format(name))
