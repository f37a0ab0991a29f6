"""Importable alias for the package directory `cubesat-apds_b200/` (a hyphen is not a valid
Python identifier): `import cubesat_apds_b200 as dunk`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("cubesat-apds_b200")
