"""Import alias for the package directory ``stablespectralelements.jl_b200/``.

The product directory carries the reference's name (with a dot, which Python cannot
import directly), so this thin alias extends its ``__path__`` to that directory:
``import sse_b200.solvers`` resolves to ``stablespectralelements.jl_b200/solvers.py``.
"""
import os as _os

_PKG_DIR = _os.path.normpath(
    _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..",
                  "stablespectralelements.jl_b200"))
__path__.append(_PKG_DIR)
PACKAGE_DIR = _PKG_DIR
