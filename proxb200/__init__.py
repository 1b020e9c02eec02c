"""proxb200 -- importable name of the package that lives in `proximalalgorithms.jl_b200/`.

The directory name required by the project layout contains a dot and cannot be imported directly, so this shim
extends its own search path with that directory: `proxb200.algorithms` is `proximalalgorithms.jl_b200/algorithms.py`.
"""
import os as _os

_IMPL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "proximalalgorithms.jl_b200")
if not _os.path.isdir(_IMPL):
    raise ImportError(f"{_IMPL} is missing")
__path__.append(_IMPL)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
