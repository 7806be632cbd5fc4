"""The C-ABI library builds for sm_100a, loads on a CPU-only box, and exports every symbol
include/fractencode_b200.h declares.  No compute calls here (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import fractencode_b200 as fb
    if not os.path.exists(fb.library_path()):
        fb.build_library()
    return fb.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fractencode_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fe_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "include/fractencode_b200.h declares %s but the library does not export it" % n


def test_python_binding_covers_header():
    import fractencode_b200.capi as capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_abi_version_and_struct_sizes(lib):
    import fractencode_b200 as fb
    assert lib.fe_abi_version() == 3
    assert fb.GRID_ITEM.itemsize == 20      # sizeof(Frac2::UniformGridItem), gpu/opencl/common.hpp:18
    assert fb.ENCODE_ITEM.itemsize == 64    # sizeof(Frac::encode_item_t)
    assert ctypes.sizeof(fb.Params) == 32


def test_no_cpu_fallback(lib):
    """Without a GPU fe_create must fail loudly (FE_ERR_NO_DEVICE), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fractencode_b200 as fb
    with pytest.raises(fb.FractencodeError) as e:
        fb.Context(0)
    assert e.value.code == -4


def test_product_does_not_touch_oracle():
    """oracle/ is test infrastructure: nothing under fractencode_b200/ or include/ may reference it."""
    bad = []
    for base in ("fractencode_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"(import|from)\s+oracle|oracle/|frac_oracle|pyoracle|libfracref", txt) and "oracle/frac_oracle.c" not in txt:
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_uniform_grid_host_mirror(fo):
    import numpy as np
    import fractencode_b200 as fb
    for (W, H, size, step) in [(512, 512, 16, 8), (64, 32, 8, 4), (8, 8, 4, 2), (240, 240, 12, 6)]:
        a, b = fb.uniform_grid(W, H, size, step), fo.uniform_grid(W, H, size, step)
        assert a.tobytes() == np.ascontiguousarray(b).tobytes()
