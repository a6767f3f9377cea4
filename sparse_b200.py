"""Import alias: ``import sparse_b200`` -> the package in ``opensearch-sparse-model-tuning-sample_b200/``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("opensearch-sparse-model-tuning-sample_b200")
sys.modules[__name__] = _pkg
