"""waiwera_b200 -- B200-native Newton-step engine for Waiwera (hand-written CUDA behind a C ABI).

Importing the package does not load the CUDA library; the first call into `waiwera_b200._lib.lib()`
does and raises if it has not been built.
"""
from . import mesh  # noqa: F401

__all__ = ["mesh", "flow", "build"]
