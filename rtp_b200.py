"""Import shim: `import rtp_b200` gives the package in `raytracing-potato_b200/` (whose directory
name is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("raytracing-potato_b200")
sys.modules[__name__] = _pkg
