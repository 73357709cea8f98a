"""`qugar.cpp` served by qugar_b200.cpp (same names, arguments, dtypes and shapes)."""
from qugar_b200.cpp import *  # noqa: F401,F403
from qugar_b200.cpp import __version__, __git_commit_hash__  # noqa: F401
from qugar_b200 import cpp as _impl

# every public name of the implementation, including the classes the reference's scripts use in type aliases
for _n in dir(_impl):
    if not _n.startswith("_"):
        globals()[_n] = getattr(_impl, _n)


def _reparam_placeholder(name):
    def _init(self, *a, **k):
        raise NotImplementedError(f"{name}: reparameterization (python/nanobind_wrappers/reparam.cpp) is not part of qugar_b200")
    return type(name, (), {"__init__": _init, "__doc__": "placeholder: lets scripts that only NAME the type (annotations) import"})


# ReparamMesh_<dim>_<range> (python/nanobind_wrappers/reparam.cpp): named by demo_div_thm_no_fenicsx.py:65-70 in a type alias
for _n in ("ReparamMesh_1_2", "ReparamMesh_2_2", "ReparamMesh_2_3", "ReparamMesh_3_3"):
    globals()[_n] = _reparam_placeholder(_n)
