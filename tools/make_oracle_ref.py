#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference modules of the hot path, staged so they can travel to the GPU box.

The reference is pure Python (no native code), so "building" it is copying the three hot-path files
byte for byte from where they lie under /root/reference into ``oracle/_ref/`` (git-ignored: never part of the history;
NOT gpurun-ignored: it ships with the tree like the built .so).  Nothing under ``oracle/_ref`` is ever edited.

    python tools/make_oracle_ref.py        # (re)creates oracle/_ref/ ; no-op where /root/reference is absent

Consumers (checker / baseline only, never the product path):
  * bench.py --impl reference and the cpu_baseline leg time THIS code (cpu_baseline.kind = "reference")
  * tests/test_dropin_callsite_gpu.py runs the reference's own MolKGNNNet with only MolGCN swapped for molkgnn_b200.MolGCN

``load()`` imports the staged modules through the test-only torch_geometric stand-in (tests/stubs) and returns them.
"""
import hashlib
import importlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["models/MolKGNN/kernels.py", "models/MolKGNN/KernelLayer.py", "models/MolKGNN/MolKGNNNet.py"]


def make(verbose=False):
    """-> True if oracle/_ref is present (freshly staged or already there), False if it cannot be made here."""
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    # package markers only (empty files): the reference itself is a flat script repo imported from its root
    for pkg in ("models", "models/MolKGNN"):
        open(os.path.join(DST, pkg, "__init__.py"), "a").close()
    json.dump({"source": REF, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("oracle/_ref staged:", ", ".join(FILES))
    return True


def available():
    return os.path.exists(os.path.join(DST, "MANIFEST.json"))


def load():
    """Import the staged reference (through tests/stubs/torch_geometric) -> dict of its modules.  The staged tree is put in
    FRONT of sys.path under the reference's own top-level package name ``models``."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (python tools/make_oracle_ref.py, needs /root/reference)")
    stubs = os.path.join(ROOT, "tests", "stubs")
    for p in (stubs, DST):
        if p not in sys.path:
            sys.path.insert(0, p)
    mods = {}
    for name in ("kernels", "KernelLayer", "MolKGNNNet"):
        mods[name] = importlib.import_module("models.MolKGNN." + name)
    f = os.path.realpath(mods["kernels"].__file__)
    if not (f.startswith(os.path.realpath(DST)) or f.startswith(REF)):
        raise RuntimeError(f"models.MolKGNN resolved to {f}, not to the staged reference")
    return mods


if __name__ == "__main__":
    ok = make(verbose=True)
    print("available" if ok else "unavailable: /root/reference is absent and oracle/_ref was never staged")
