"""Importable alias for the ``g-nerf_b200`` package (a hyphen is not a valid identifier)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module('g-nerf_b200')
sys.modules[__name__] = _pkg
