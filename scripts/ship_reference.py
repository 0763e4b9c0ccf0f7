#!/usr/bin/env python
"""Place the UNMODIFIED reference where the GPU box can see it: /root/reference -> baseline/_ref/.

`baseline/_ref/` is git-ignored (the reference's sources never enter this repository's history) but NOT
gpurun-ignored, so it travels with the snapshot like the built .so.  The reference has no setup.py /
pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to install; this script copies
the Python sources and `hyperparam.ini` verbatim instead (no edits -- `tests/test_reference_shipping.py` checks the
copies byte for byte against /root/reference when that exists).

Used by: `bench.py --impl reference` / `--impl reference-cuda` (the reference's own Trainer on the host cores / on
the B200 through stock PyTorch eager) and `tests/test_main_gpu.py` (the reference's `main.py` driving this
repository's `disvae` package).  Run in the build container:  python scripts/ship_reference.py
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DISVAE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
ITEMS = ["disvae", "utils", "main.py", "main_viz.py", "hyperparam.ini", "LICENSE"]


def ship(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("no reference at %s: keeping whatever baseline/_ref holds" % REF)
        return os.path.isdir(os.path.join(DST, "disvae"))
    os.makedirs(DST, exist_ok=True)
    for it in ITEMS:
        src, dst = os.path.join(REF, it), os.path.join(DST, it)
        if os.path.isdir(src):
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copyfile(src, dst)
    if verbose:
        print("reference shipped to", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if ship() else 1)
