"""Importable alias of the package directory `neural-volume-super-resolution_b200/` (hyphens are not
valid in an import statement).  `import nvsr_b200` gives the package itself."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("neural-volume-super-resolution_b200")
sys.modules[__name__] = _pkg
