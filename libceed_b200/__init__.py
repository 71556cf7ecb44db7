"""ceed-b200: a B200-native (sm_100a) libCEED operator-apply backend.

The product is the C-ABI shared library `libceed_b200/lib/libceed_b200.so` (include/ceed_b200.h) and the libCEED backend
plugin that binds it to resource "/gpu/cuda/b200" (libceed_b200/backend/).  This Python package is a thin ctypes mirror
of the reference's object interface, used by tests and bench.py.  There is no CPU fallback.
"""
from .ceed import (BASIS_NONE, ELEMRESTRICTION_NONE, VECTOR_ACTIVE, VECTOR_NONE, Ceed, CeedError)  # noqa: F401
from ._lib import *  # noqa: F401,F403
