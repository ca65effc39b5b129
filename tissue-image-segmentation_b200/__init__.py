"""tiseg_b200 — B200-native test-time instance pipeline (post-process + evaluation) of `tiseg`.

Host-side mirror of the reference's Python API for this path over a C-ABI CUDA library
(``csrc/`` -> ``libtiseg_b200.so``, declared in ``include/tiseg_b200.h``).  There is no CPU
fallback: every compute entry point raises if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
