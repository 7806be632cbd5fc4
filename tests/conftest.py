import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fo():
    """Our C restatement of the reference path (the checker)."""
    from oracle import pyoracle
    pyoracle.build(ref=os.path.isdir("/root/reference/encode"))
    return pyoracle.restatement()


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(GOLDEN, "goldens.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def lenna():
    return np.load(os.path.join(GOLDEN, "lenna512_luma.npz"))["luma"]


@pytest.fixture(scope="session")
def images(fo, lenna):
    return {
        "lenna": lenna,
        "natural256": fo.synth_image(256, 256, 1234, 0),
        "noise256": fo.synth_image(256, 256, 1234, 1),
        "pattern256": fo.synth_image(256, 256, 1234, 2),
        "natural240": fo.synth_image(240, 240, 1234, 0),
    }


@pytest.fixture(scope="session")
def ctx():
    """A B200 context through the C ABI; fails (not skips) when the library or the GPU is missing."""
    import fractencode_b200 as fb
    c = fb.Context(0)
    yield c
    c.close()


def fnv1a64(b: bytes) -> str:
    h = 0xCBF29CE484222325
    for byte in b:
        h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
