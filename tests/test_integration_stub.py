"""INTEGRATION.md's reference-side engine stub compiles and links against the REAL reference headers and our C ABI
(needs /root/reference; the GPU box does not have it, so the prebuilt binary travels and runs there)."""
import hashlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_engine_main")


def build():
    import fractencode_b200 as fb
    if not os.path.exists(fb.library_path()):
        fb.build_library()
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=gnu++20", "-O2", "-include", "mutex", "-include", "condition_variable", "-Wno-deprecated-declarations",
           "-I", REF, "-I", os.path.join(REF, "thirdparty/gsl/include"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests/integration"), "-o", BIN, os.path.join(ROOT, "tests/integration/ref_engine_main.cpp"),
           "-L", os.path.join(ROOT, "fractencode_b200"), "-lfractencode_b200", "-Wl,-rpath,$ORIGIN/../../fractencode_b200",
           "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart", "-pthread"]
    subprocess.run(cmd, check=True)


def test_stub_compiles_against_reference_headers():
    if not os.path.isdir(os.path.join(REF, "encode")):
        pytest.skip("/root/reference not present")
    build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
@pytest.mark.parametrize("name,S,T,cls", [("lenna_16_8", 16, 8, 0), ("lenna_16_8_cls", 16, 8, 1)])
def test_reference_types_through_the_stub(lenna, goldens, tmp_path, name, S, T, cls):
    """The reference's own ImagePlane / UniformGrid / encode_item_t driven through init/encode/finalize/result."""
    if not os.path.exists(BIN):
        pytest.skip("ref_engine_main was not prebuilt (needs /root/reference at build time)")
    raw = tmp_path / "luma.raw"
    raw.write_bytes(lenna.tobytes())
    p = subprocess.run([BIN, str(raw), "512", "512", str(S), str(T), str(cls)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-500:] + p.stderr[-500:]
    # the reference engine's destructor prints "<name>, tasks done: N" (encode/EncodingEngine2.hpp:60-62): keep the item lines
    lines = sorted((l for l in p.stdout.strip().split("\n") if " | " in l), key=lambda l: (int(l.split()[1]), int(l.split()[0])))
    md5 = hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest()
    # the stub is compiled without FMA contraction flags -> FRAC_FMA_BUILD = 0 -> the non-FMA golden
    assert md5 == goldens[name]["md5_nofma"]
