"""Shim package: put `qugar_b200/shim` on sys.path and the reference's FEniCSx-free scripts (`import qugar`,
`import qugar.cpp`, e.g. python/demo/demo_div_thm_no_fenicsx.py) run unmodified on libqugar_b200.so.

Only `qugar.cpp` exists here -- the nanobind module of the reference (python/nanobind_wrappers/qugar.cpp:35-60).  The DOLFINx
layers of the real package (`qugar.mesh`, `qugar.dolfinx`, ...) are outside the accelerated path and are not shimmed."""
from . import cpp  # noqa: F401

__version__ = cpp.__version__
has_FEniCSx = False
