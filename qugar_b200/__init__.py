"""qugar_b200 -- B200-native (sm_100a) engine for QUGaR's implicit-domain cut-cell quadrature path.

`qugar_b200.cpp` mirrors the reference's nanobind module `qugar.cpp` for that path; the arithmetic runs in
hand-written CUDA kernels behind the C ABI declared in include/qugar_b200.h (libqugar_b200.so).
"""
from . import cpp  # noqa: F401

__all__ = ["cpp"]
