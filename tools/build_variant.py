#!/usr/bin/env python
"""Build a variant of libofr_b200.so with extra -D flags (kernel A/B experiments on the GPU box).

    python tools/build_variant.py NAME -DOFR_STREAM_BM=8 ...   ->  <package>/build/variants/libofr_NAME.so
    OFR_B200_LIB=<that path> python tools/opbench.py ...          (the Python loader's development switch)
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ofr_b200  # noqa: E402

name = sys.argv[1]
defines = [a[2:] for a in sys.argv[2:] if a.startswith("-D")]
out_dir = os.path.join(ROOT, ofr_b200.__name__ if os.path.isdir(os.path.join(ROOT, ofr_b200.__name__)) else
                       os.path.basename(os.path.dirname(ofr_b200._lib.__file__)), "build", "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libofr_{name}.so")
print(ofr_b200.build(force=True, defines=defines, out=out))
# the variant's object files are not needed again (and would travel to the GPU box with every snapshot)
shutil.rmtree(os.path.join(os.path.dirname(out_dir), os.path.basename(out) + ".d"), ignore_errors=True)
