"""B200-native (sm_100a) neural-sparse encoding hot path.

Drop-in for the hot path of zhichao-aws/opensearch-sparse-model-tuning-sample: the package mirrors the
reference's ``scripts`` module layout (``scripts.model.sparse_encoders``, ``scripts.train.loss`` ...), with the
arithmetic running in hand-written CUDA kernels behind the C ABI of ``libsparse_b200.so`` (include/sparse_b200.h).
The directory name contains hyphens, so import it through the ``sparse_b200`` alias module at the repository
root (``import sparse_b200``) or ``importlib.import_module("opensearch-sparse-model-tuning-sample_b200")``.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
