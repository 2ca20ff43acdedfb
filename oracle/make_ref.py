"""Stages the UNMODIFIED reference's Python sources under oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python (no build system, no setup.py), so the
base contract's `pip install --target baseline/_ref /root/reference` has nothing to install; this recipe is
its equivalent: a byte-for-byte copy of the files the per-step path imports, made in the build container
where /root/reference exists, kept out of git history, and shipped to the GPU box with the working tree
(like the built .so files).  Consumers: `bench.py --impl reference` / `--impl pytorch-gpu` (the reference
arm times the reference's own `Trainer.process_batch`) and the drop-in tests, which run the unchanged
`trainer.py` against `fusiondepth_b200/dropin`.  Nothing under fusiondepth_b200/ reads it.

    python oracle/make_ref.py            # copy /root/reference -> oracle/_ref (no-op when absent)
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FD_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
DIRS = ("networks", "datasets")


def stage(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds the reference sources afterwards."""
    if not os.path.isdir(SRC):
        return os.path.isfile(os.path.join(DST, "trainer.py"))
    os.makedirs(DST, exist_ok=True)
    n = 0
    for f in sorted(os.listdir(SRC)):
        if f.endswith(".py"):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
            n += 1
    for d in DIRS:
        dst = os.path.join(DST, d)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, d), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for lic in ("LICENSE",):
        if os.path.exists(os.path.join(SRC, lic)):
            shutil.copyfile(os.path.join(SRC, lic), os.path.join(DST, lic))
    # the copy must be the unmodified reference
    for f in ("trainer.py", "refiner.py", "layers.py", "kitti_utils.py", "networks/depth_decoder.py"):
        assert filecmp.cmp(os.path.join(SRC, f), os.path.join(DST, f), shallow=False), f
    if verbose:
        print("oracle/_ref: staged %d top-level modules + %s from %s" % (n, ", ".join(DIRS), SRC))
    return True


def root() -> str | None:
    """Where the reference can be imported from on this machine: /root/reference in the build container,
    else the staged copy; None when neither exists."""
    if os.path.isfile(os.path.join(SRC, "trainer.py")):
        return SRC
    if os.path.isfile(os.path.join(DST, "trainer.py")):
        return DST
    return None


if __name__ == "__main__":
    ok = stage()
    sys.exit(0 if ok else 1)
