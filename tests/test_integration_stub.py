"""INTEGRATION.md's reference-side engine stub compiles and links against the REAL reference headers and our C ABI
(needs /root/reference; the GPU box does not have it, so the prebuilt binary travels and runs there)."""
import hashlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_engine_main")
CORE_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_core_main")


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
    # the engine registered in the reference's EncodingEngineCore2: Encoder2 + the reference's own sources, except that
    # EngineCoreWithB200.cpp stands in for encode/EncodingEngine2.cpp (the constructor with the engine slot filled)
    ref_srcs = [os.path.join(REF, f) for f in ("image/ImageStatistics.cpp", "encode/Classifier2.cpp", "image/ImageIO.cpp", "image/transform.cpp")]
    stb = os.path.join(ROOT, "oracle", "_ref", "stb_impl.o")
    if not os.path.exists(stb):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    core = list(cmd)
    i = core.index("-o")
    core[i + 1] = CORE_BIN
    core[i + 2: i + 3] = [os.path.join(ROOT, "tests/integration/ref_core_main.cpp"), os.path.join(ROOT, "tests/integration/EngineCoreWithB200.cpp")] + ref_srcs + [stb]
    subprocess.run(core, check=True)


def _dump_md5(stdout):
    # the reference engines' destructors print "<name>, tasks done: N" (encode/EncodingEngine2.hpp:60-62): keep the item lines
    lines = sorted((l for l in stdout.strip().split("\n") if " | " in l), key=lambda l: (int(l.split()[1]), int(l.split()[0])))
    return hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest(), len(lines)


def test_stub_compiles_against_reference_headers():
    if not os.path.isdir(os.path.join(REF, "encode")):
        pytest.skip("/root/reference not present")
    build()
    assert os.path.exists(BIN) and os.path.exists(CORE_BIN)


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
    md5, _ = _dump_md5(p.stdout)
    # the stub is compiled without FMA contraction flags -> FRAC_FMA_BUILD = 0 -> the non-FMA golden
    assert md5 == goldens[name]["md5_nofma"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,S,T,cls,nocpu", [("lenna_16_8", 16, 8, 0, 1), ("lenna_16_8_cls", 16, 8, 1, 1), ("lenna_16_8_cls", 16, 8, 1, 0)])
def test_engine_registered_in_the_reference_engine_core(lenna, goldens, tmp_path, name, S, T, cls, nocpu):
    """The B200 engine as a member of EncodingEngineCore2::_engines (the slot of encode/EncodingEngine2.cpp:21-26), fed by the
    reference's own job queue through Encoder2 -- alone (--nocpu semantics) and next to the CPU engines.  The reference's
    wait loop can lose a wake-up: watchdog + retry."""
    if not os.path.exists(CORE_BIN):
        pytest.skip("ref_core_main was not prebuilt (needs /root/reference at build time)")
    raw = tmp_path / "luma.raw"
    raw.write_bytes(lenna.tobytes())
    out = None
    for attempt in range(6):
        try:
            p = subprocess.run([CORE_BIN, str(raw), "512", "512", str(S), str(T), str(cls), str(nocpu)], capture_output=True, text=True, timeout=120)
        except subprocess.TimeoutExpired:
            continue                                  # the lost wake-up of EncodingEngineCore2::encode
        assert p.returncode == 0, p.stdout[-500:] + p.stderr[-500:]
        out = p.stdout
        break
    assert out is not None, "EncodingEngineCore2::encode hung six times in a row"
    md5, n = _dump_md5(out)
    assert n == (512 // T) ** 2
    assert "B200, tasks done:" in out
    assert md5 == goldens[name]["md5_nofma"]
