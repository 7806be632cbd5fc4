"""The C++ drop-in headers (fractencode_b200/host): a main.cpp-style client written against the
reference's include paths and class names, run on the GPU and compared with the real reference's output."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "host", "dropin_main")


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "fractencode_b200", "csrc")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "fractencode_b200", "host")], check=True)


def test_dropin_client_compiles_against_reference_include_paths():
    """CPU check: the client uses only "encode/...", "image/..." includes and Frac/Frac2 names."""
    build()
    assert os.path.exists(BIN)
    src = open(os.path.join(ROOT, "tests", "host", "dropin_main.cpp")).read()
    incs = re.findall(r'#include "([^"]+)"', src)
    assert incs and all(i.split("/")[0] in ("encode", "image", "utils") for i in incs), incs
    ref = "/root/reference"
    if os.path.isdir(ref):  # every forwarding header shadows a file that exists in the reference
        host = os.path.join(ROOT, "fractencode_b200", "host")
        for d in ("encode", "image", "utils"):
            for fn in os.listdir(os.path.join(host, d)):
                assert os.path.exists(os.path.join(ref, d, fn)), "%s/%s is not a reference header" % (d, fn)


def run(lenna, tmp_path, args):
    raw = tmp_path / "luma.raw"
    raw.write_bytes(lenna.tobytes())
    p = subprocess.run([BIN, str(raw), "512", "512"] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-500:] + p.stderr[-500:]
    info = dict(kv.split("=") for kv in p.stderr.strip().split("\n")[-1].split())
    return hashlib.md5(p.stdout.encode()).hexdigest(), info


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [
    ("lenna_16_8", [16, 8, 1, 0.0, -1.0]),          # BASELINE config 1
    ("lenna_16_8_cls", [16, 8, 0, 0.0, -1.0]),
    ("lenna_16_4", [16, 4, 1, 0.0, -1.0]),          # reference default geometry
    ("lenna_qt_16_4_cls_thr5", [32, 16, 0, 5.0, -1.0, "Q"]),  # BASELINE config 2
])
def test_dropin_client_matches_reference(lenna, goldens, tmp_path, name, args):
    build()
    c = goldens[name]
    for fma in (0, 1):
        a = list(args)
        quad = a and a[-1] == "Q"
        if quad:
            a = a[:-1]
        a = a + [fma] + (["quadtree", c["qt"][0], c["qt"][1]] if quad else [])
        md5, info = run(lenna, tmp_path, a)
        assert md5 == (c["md5_fma"] if fma else c["md5_nofma"]), (name, fma)
        assert int(info["items"]) == c["n_items"]
        if not fma:
            assert int(info["decode_iterations"]) == c["decode_iterations"]
            assert float(info["decode_rms"]) == c["decode_rms"]
            assert info["decode_fnv1a64"] == c["decode_fnv1a64"]
