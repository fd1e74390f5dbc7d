"""Stage the UNMODIFIED reference where the GPU box can reach it: /root/reference -> oracle/_ref/ (git-ignored, shipped
by gpurun).  TEST / MEASUREMENT INFRASTRUCTURE: used by `bench.py --impl reference` (the reference's own CPU PyTorch path
as the baseline arm) and by tests/test_gpu_dropin.py (the reference's real driver functions running
on the univl_b200 package).  The reference is a pure-Python repo with no build step (no setup.py / pyproject.toml):
"building" it is a file copy.  Nothing from oracle/_ref is imported by the product package.

    python oracle/build_ref.py            # no-op (exit 0) when /root/reference is absent and oracle/_ref exists
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("UNIVL_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP = ("modules", "dataloaders", "main_pretrain.py", "main_task_caption.py", "main_task_retrieval.py", "metrics.py",
        "util.py", "LICENSE", "NOTICE.md")


def build(force=False):
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST) and not force:
        same = all(os.path.exists(os.path.join(DST, k)) for k in KEEP if os.path.exists(os.path.join(SRC, k)))
        if same:
            return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for k in KEEP:
        s, d = os.path.join(SRC, k), os.path.join(DST, k)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.exists(s):
            shutil.copy2(s, d)
    return DST


def ref_root():
    """oracle/_ref when staged, else the authoring container's /root/reference, else None"""
    if os.path.isdir(os.path.join(DST, "modules")):
        return DST
    if os.path.isdir(os.path.join(SRC, "modules")):
        return SRC
    return None


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "reference not available (no %s, no %s)" % (SRC, DST))
