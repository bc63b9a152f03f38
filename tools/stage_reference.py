#!/usr/bin/env python
"""Stage the UNMODIFIED reference model files into ``baseline/_ref/`` so they travel to the GPU box.

``/root/reference`` exists only in the dev container; ``baseline/_ref/`` is git-ignored (never part of
the history, outside the copy checks) but not gpurun-ignored, so the snapshot that goes to the B200 box
carries it.  The GPU tests (``tests/test_gpu_reference_models.py``) and ``bench.py``'s FlowNet2 workload
import the reference's own classes from there, with ``ofr_b200.install_shims()`` serving the three
custom-layer import paths the reference tree lacks.

The reference has no ``setup.py`` / ``pyproject.toml`` (so the base contract's ``pip install --target
baseline/_ref`` has nothing to install): staging is a byte-for-byte copy of the Python files, recorded
with their SHA-256 in ``baseline/_ref/MANIFEST.json``.

    python tools/stage_reference.py            # copy (idempotent)
    python tools/stage_reference.py --check    # verify the staged files still equal the reference
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

# model / loss code on or around the hot path (SURVEY.md section 8b "what calls it"); data pipeline, CLIs and
# metrics are out of scope and stay behind
FILES = [
    "loss.py", "models.py", "utils.py",
    "PWC/__init__.py", "PWC/models/__init__.py", "PWC/models/PWCNet.py",
    "flownet2/__init__.py", "flownet2/models.py",
    "flownet2/networks/__init__.py", "flownet2/networks/FlowNetC.py", "flownet2/networks/FlowNetS.py",
    "flownet2/networks/FlowNetSD.py", "flownet2/networks/FlowNetFusion.py", "flownet2/networks/submodules.py",
    "FlowNetS/__init__.py", "FlowNetS/FlowNetS.py", "FlowNetS/FlowNetC.py", "FlowNetS/util.py",
]


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(verbose=False):
    """Copy FILES from the reference tree; returns the staging directory (or None when there is no reference)."""
    if not os.path.isdir(REF):
        return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(src) != _sha(dst):
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", rel)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    return DST


def check():
    man = json.load(open(os.path.join(DST, "MANIFEST.json")))
    bad = [rel for rel, h in man["files"].items() if _sha(os.path.join(DST, rel)) != h]
    if os.path.isdir(REF):
        bad += [rel for rel in man["files"] if _sha(os.path.join(REF, rel)) != man["files"][rel]]
    return sorted(set(bad))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.check:
        bad = check()
        print("modified:", bad if bad else "none")
        sys.exit(1 if bad else 0)
    print(stage(verbose=True))
