"""The oracle restatement against the committed outputs of the REAL reference (tests/golden,
made by tests/golden/make_golden.py from oracle/_ref)."""
import hashlib
import os

import numpy as np
import pytest

from tests.cases import assert_items_equal, md5_dump, run_case_oracle
from tests.conftest import GOLDEN, fnv1a64

# cases cheap enough for the CPU suite (the rest are covered by the -m gpu run, which also calls the oracle)
FAST = ["lenna_16_8_cls", "lenna_32_16", "lenna_64_32", "lenna_64_32_cls", "lenna_qt_32_8_thr10", "nat256_16_8",
        "nat256_8_4_cls", "nat256_32_16", "nat256_64_32", "nat256_qt_32_4_cls_thr20_smax1", "noise256_16_8",
        "noise256_32_16", "noise256_64_32_cls", "pattern256_16_8", "pattern256_8_4_cls", "pattern256_16_8_thr3",
        "pattern256_qt_16_4_thr1", "nat240_12_6", "nat240_24_12_cls_thr30", "nat256_16_4", "nat256_32_8_cls"]


def test_fixture_image_hashes(images, goldens):
    for name, c in goldens.items():
        assert hashlib.md5(images[c["image"]].tobytes()).hexdigest() == c["image_md5"], name


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_golden(fo, images, goldens, name):
    c = goldens[name]
    img = images[c["image"]]
    items, counts = run_case_oracle(fo, img, c, fma=False)
    assert len(items) == c["n_items"]
    assert md5_dump(items) == c["md5_nofma"]
    if counts is not None:
        assert counts == c["level_counts"]
    items_f, _ = run_case_oracle(fo, img, c, fma=True)
    assert md5_dump(items_f) == c["md5_fma"]
    path = os.path.join(GOLDEN, "items_%s.npz" % name)
    if os.path.exists(path):
        z = np.load(path)
        assert_items_equal(items, z["items"], name)
        assert_items_equal(items_f, z["items_fma"], name + " (fma)")
    dec, it, rms = fo.decode(items, img.shape[1], img.shape[0])
    assert it == c["decode_iterations"] and rms == c["decode_rms"]
    assert fnv1a64(dec.tobytes()) == c["decode_fnv1a64"]


def test_oracle_vs_live_reference_random(fo):
    """Seeded random blocky/noisy images, thresholds, sMax, both FMA modes: restatement == compiled reference."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(7)
    libs = [(po.reference(False), False), (po.reference(True), True)]
    if libs[0][0] is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for trial in range(6):
        W = H = 64
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
        if trial % 2:
            img = np.kron(rng.integers(0, 256, (H // 8, W // 8), dtype=np.uint8), np.ones((8, 8), np.uint8))
            img = (img.astype(int) + rng.integers(0, 3, (H, W))).clip(0, 255).astype(np.uint8)
        S, T = [(8, 4), (16, 8), (16, 4), (32, 16), (8, 2), (4, 2)][trial]
        for ref, fma in libs:
            p = fo.params(thr=[0.0, 40.0, 300.0][trial % 3], smax=[-1.0, 1.0][trial % 2], classifier=bool(trial & 2), fma=fma)
            dom, rg = fo.uniform_grid(W, H, S, S // 2), fo.uniform_grid(W, H, T, T)
            if trial & 2:
                dom, rg = fo.preclassify(img, dom), fo.preclassify(img, rg)
            a = fo.encode_level(img, img, dom, rg, p)
            b = ref.encode_level(img, img, dom, rg, p)
            assert_items_equal(a, b, "trial %d fma=%s" % (trial, fma))
