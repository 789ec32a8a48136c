"""Import alias: ``import vbfem_b200`` loads the package whose directory name
(variational-bayesian-inference-for-computational-mechanics_b200) is not a
valid Python identifier."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("variational-bayesian-inference-for-computational-mechanics_b200")
sys.modules[__name__] = _pkg
