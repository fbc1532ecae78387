"""gtok - B200-native graph tokenization engine (host side).

Importing this package loads libgtok.so (hand-written sm_100a kernels behind the
C-ABI of include/gtok.h) and raises if it is missing: there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA extension is not built)
from .batch import GraphBatch  # noqa: F401
