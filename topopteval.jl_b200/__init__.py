"""topopteval.jl_b200 — B200-native strain-energy evaluation path of jezekon/TopOptEval.jl.

The directory name contains a dot, so it is imported through `__graft_entry__.load_package()` (registers it as
`topopteval_jl_b200`).  Contents: `csrc/` (CUDA kernels + C ABI → libtopopt_b200.so), `_lib.py` (ctypes binding),
`api.py` (mirror of the reference's Julia API), `julia/` (ccall shim), `vtu.py` / `meshgen.py` (harness I/O).
"""
from . import _lib, meshgen, parallel, vtu  # noqa: F401
from ._lib import Context, TopOptError  # noqa: F401
from .api import *  # noqa: F401,F403
from . import api  # noqa: F401
