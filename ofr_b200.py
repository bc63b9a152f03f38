"""Short importable alias of the product package.

The package directory is named after the reference repository
(``self-supervised-medical-image-registration-using-deep-optical-flow-estimation-with-brain-mri-data_b200``),
which is not a valid Python identifier; ``import ofr_b200`` loads it through importlib and
re-exports it under this name (sub-modules included: ``ofr_b200.layers``, ``ofr_b200.harness`` ...).
"""
import importlib
import os
import sys

PACKAGE_NAME = "self-supervised-medical-image-registration-using-deep-optical-flow-estimation-with-brain-mri-data_b200"

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)

_pkg = importlib.import_module(PACKAGE_NAME)
_alias = __name__
sys.modules[_alias] = _pkg
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(PACKAGE_NAME + "."):
        sys.modules[_alias + _name[len(PACKAGE_NAME):]] = _mod
