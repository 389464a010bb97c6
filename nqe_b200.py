"""Importable alias of the `naive-query-engine_b200` package (hyphenated directory)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("naive-query-engine_b200")
sys.modules[__name__] = _pkg
