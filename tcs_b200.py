"""Importable alias of the `temporally-consistent-stereo-matching_b200` package (hyphens are not valid in
an import statement).  `import tcs_b200` == that package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("temporally-consistent-stereo-matching_b200")
sys.modules[__name__] = _pkg
