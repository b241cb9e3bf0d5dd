"""``import vit2spn`` → the package in ``vit-2spn_b200/`` (hyphenated directory name)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("vit-2spn_b200")
sys.modules[__name__] = _pkg
