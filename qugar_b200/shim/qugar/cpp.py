"""`qugar.cpp` served by qugar_b200.cpp (same names, arguments, dtypes and shapes)."""
from qugar_b200.cpp import *  # noqa: F401,F403
from qugar_b200.cpp import __version__, __git_commit_hash__  # noqa: F401
from qugar_b200 import cpp as _impl

# every public name of the implementation, including the classes the reference's scripts use in type aliases
for _n in dir(_impl):
    if not _n.startswith("_"):
        globals()[_n] = getattr(_impl, _n)
