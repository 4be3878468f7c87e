"""Import alias: the package directory name contains hyphens (it mirrors the reference's
repository name), which ``import`` cannot spell.  ``import udal_b200`` gives the same module."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("uncertainty-detection-autolabeling_b200")
sys.modules[__name__] = _pkg
