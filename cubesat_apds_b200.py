"""Importable alias for the package directory `cubesat-apds_b200/` (a hyphen is not a valid
Python identifier): `import cubesat_apds_b200 as dunk`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("cubesat-apds_b200")
# alias the package AND its sub-modules: otherwise `from cubesat_apds_b200._lib import x` would import
# `_lib` a second time under the alias (a second DunkError class, a second library handle)
for _name, _mod in list(sys.modules.items()):
    if _name == "cubesat-apds_b200" or _name.startswith("cubesat-apds_b200."):
        sys.modules[__name__ + _name[len("cubesat-apds_b200"):]] = _mod
sys.modules[__name__] = _pkg
