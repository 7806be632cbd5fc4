"""fractencode_b200 -- B200-native fractal-encoding search behind the reference's encode/ API.

The product is the C-ABI shared library (include/fractencode_b200.h, built from
fractencode_b200/csrc for sm_100a) plus the C++ drop-in headers under
fractencode_b200/host/.  This Python package is a thin ctypes front-end over the same
C ABI, used by bench.py and the tests.  There is no CPU fallback: importing works
anywhere, creating a Context needs a B200 and the built library.
"""
from .capi import (  # noqa: F401
    ENCODE_ITEM,
    GRID_ITEM,
    Context,
    FractencodeError,
    Params,
    Stats,
    build_library,
    library_path,
    load_library,
    uniform_grid,
)

__all__ = ["Context", "Params", "Stats", "FractencodeError", "GRID_ITEM", "ENCODE_ITEM", "uniform_grid",
           "load_library", "library_path", "build_library"]
