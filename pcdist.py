"""Importable alias: `import pcdist` == the package in ./3dpointcloudattack_b200/ (whose
directory name is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("3dpointcloudattack_b200")
