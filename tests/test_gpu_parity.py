"""Parity of the CUDA path (through the C ABI) with the oracle and with the committed outputs of the
real reference.  Bit-exact: domain, isometry, distance, contrast, brightness (both FMA modes)."""
import os

import numpy as np
import pytest

from tests.cases import assert_items_equal, md5_dump, run_case_gpu, run_case_oracle
from tests.conftest import GOLDEN, fnv1a64
from tests.test_oracle_kat import (CLASSIFIER_KAT, EST_EXPECTED, EST_SRC, EST_TGT, MATCH_SRC, MATCH_TGT, R270)

pytestmark = pytest.mark.gpu

ALL_CASES = ["lenna_16_8", "lenna_16_8_cls", "lenna_16_4", "lenna_8_4_cls", "lenna_32_16", "lenna_64_32", "lenna_64_32_cls",
             "lenna_qt_16_4_cls_thr5", "lenna_qt_16_4_cls_thr2", "lenna_qt_32_8_thr10", "nat256_16_8", "nat256_8_4_cls",
             "nat256_32_16", "nat256_64_32", "nat256_qt_32_4_thr8", "nat256_qt_32_4_cls_thr20_smax1", "noise256_16_8",
             "noise256_32_16", "noise256_64_32_cls", "pattern256_16_8", "pattern256_8_4_cls", "pattern256_16_8_thr3",
             "pattern256_qt_16_4_thr1", "nat240_12_6", "nat240_24_12_cls_thr30", "nat256_16_4", "nat256_32_8_cls"]


def test_transform_matcher_kat_gpu(ctx):
    """tests/TransformMatcherTest.cpp:9-36 through fe_set_images + fe_encode_level."""
    import fractencode_b200 as fb
    ctx.set_images(MATCH_SRC, MATCH_TGT)
    dom = np.array([(0, 0, 4, 4, -1)], fb.GRID_ITEM)
    rng = np.array([(0, 0, 2, 2, -1)], fb.GRID_ITEM)
    out = ctx.encode_level(dom, rng, fb.Params(0.0, 100.0))
    assert out[0]["distance"] == 0.0 and out[0]["transform"] == R270
    assert out[0]["contrast"] < 1.0 and out[0]["brightness"] < 1.0


def test_transform_estimator_kat_gpu(ctx, fo):
    """tests/TransformEstimatorTest.cpp:10-48 (separate source/target planes, 2x2 ranges)."""
    import fractencode_b200 as fb
    ctx.set_images(EST_SRC, EST_TGT)
    dom, rng = fb.uniform_grid(8, 8, 4, 2), fb.uniform_grid(4, 4, 2, 2)
    out = ctx.encode_level(dom, rng, fb.Params(0.0, 100.0))
    got = {(int(e["x"]), int(e["y"])): (int(e["match_x"]), int(e["match_y"])) for e in out}
    assert got == EST_EXPECTED
    want = fo.encode_level(EST_SRC, EST_TGT, fo.uniform_grid(8, 8, 4, 2), fo.uniform_grid(4, 4, 2, 2), fo.params(0.0, 100.0))
    assert_items_equal(out, want, "estimator KAT")


def test_classifier_kat_gpu(ctx, lenna, fo):
    """tests/ClassifierTest.cpp:24-41 through fe_classify, then every block of three grids vs the oracle."""
    import fractencode_b200 as fb
    ctx.set_image(lenna)
    for size, rows in CLASSIFIER_KAT.items():
        items = np.array([(x, y, size, size, -1) for x, y, _ in rows], fb.GRID_ITEM)
        assert ctx.classify(items).tolist() == [c for _, _, c in rows]
    for size, step in [(4, 4), (16, 8), (64, 32), (2, 2)]:
        grid = fb.uniform_grid(512, 512, size, step)
        want = fo.preclassify(lenna, grid)["bin"]
        assert (ctx.classify(grid) == want).all(), size


def test_synthetic_generator_matches_oracle(ctx, fo):
    for kind in (0, 1, 2):
        ctx.set_synthetic_image(256, 128, 99, kind)
        assert (ctx.get_image() == fo.synth_image(256, 128, 99, kind)).all()


@pytest.mark.parametrize("name", ALL_CASES)
def test_golden_case(ctx, fo, images, goldens, name):
    """Transform lists equal to the real reference's (md5 of the SURVEY-8c dump, both FMA builds),
    full items where committed, then decode through fe_decode."""
    c = goldens[name]
    img = images[c["image"]]
    items, counts = run_case_gpu(ctx, img, c, fma=False)
    path = os.path.join(GOLDEN, "items_%s.npz" % name)
    if os.path.exists(path):
        assert_items_equal(items, np.load(path)["items"], name)
    assert len(items) == c["n_items"]
    assert md5_dump(items) == c["md5_nofma"], name
    if counts is not None:
        assert counts == c["level_counts"]
    items_f, _ = run_case_gpu(ctx, img, c, fma=True)
    assert md5_dump(items_f) == c["md5_fma"], name + " (fma)"
    dec, it, rms = ctx.decode(items, img.shape[1], img.shape[0])
    assert (it, rms) == (c["decode_iterations"], c["decode_rms"])
    assert fnv1a64(dec.tobytes()) == c["decode_fnv1a64"]


def test_random_lists_vs_oracle(ctx, fo):
    """Arbitrary (non-lattice) domain/range lists incl. odd origins (generic geometry), padded stride,
    thresholds, sMax, classifier with caller-provided bins."""
    import fractencode_b200 as fb
    rs = np.random.default_rng(11)
    for trial in range(10):
        W, H = 96, 64
        buf = rs.integers(0, 256, (H, W + 32), dtype=np.uint8)
        if trial % 3 == 0:
            buf[:, :W] = np.kron(rs.integers(0, 256, (H // 8, W // 8), dtype=np.uint8), np.ones((8, 8), np.uint8))
        img = buf[:, :W]
        S, T = [(8, 4), (16, 8), (12, 4), (32, 16), (4, 2), (12, 6), (16, 4), (8, 4), (24, 8), (6, 2)][trial]
        nD, nR = 150, 60
        even = trial % 2 == 0
        dom = np.zeros(nD, fb.GRID_ITEM)
        dom["x"] = rs.integers(0, W - S + 1, nD) & (~1 if even else ~0)
        dom["y"] = rs.integers(0, H - S + 1, nD) & (~1 if even else ~0)
        dom["w"] = dom["h"] = S
        dom["bin"] = -1
        rng = np.zeros(nR, fb.GRID_ITEM)
        rng["x"] = rs.integers(0, W - T + 1, nR)
        rng["y"] = rs.integers(0, H - T + 1, nR)
        rng["w"] = rng["h"] = T
        rng["bin"] = -1
        cls = trial in (1, 4, 5, 8)
        thr = [0.0, 25.0, -1.0, 400.0][trial % 4]
        for fma in (False, True):
            p = fb.Params(thr, [-1.0, 0.8][trial % 2], cls, fma)
            ctx.set_image(img)
            got = ctx.encode_level(dom, rng, p)
            want = fo.encode_level(img, img, dom, rng, fo.params(thr, [-1.0, 0.8][trial % 2], cls, fma))
            assert_items_equal(got, want, "trial %d fma=%s" % (trial, fma))


def test_caller_bins_are_honoured(ctx, fo):
    """Classifier2::compare uses the stored bin unless it is -1 (Classifier2.cpp:70-81)."""
    import fractencode_b200 as fb
    img = fo.synth_image(64, 64, 5, 0)
    dom, rng = fb.uniform_grid(64, 64, 16, 8), fb.uniform_grid(64, 64, 8, 8)
    dom["bin"] = np.arange(len(dom)) % 3      # made-up classes
    rng["bin"] = np.arange(len(rng)) % 3
    rng["bin"][::5] = -1                      # these get computed like the reference does
    ctx.set_image(img)
    got = ctx.encode_level(dom, rng, fb.Params(0.0, -1.0, True))
    want = fo.encode_level(img, img, dom, rng, fo.params(0.0, -1.0, True))
    assert_items_equal(got, want)


def test_empty_bucket_gives_default_item(ctx, fo):
    import fractencode_b200 as fb
    img = fo.synth_image(64, 64, 6, 0)
    dom, rng = fb.uniform_grid(64, 64, 16, 8), fb.uniform_grid(64, 64, 8, 8)
    dom["bin"] = 2
    rng["bin"] = 4
    ctx.set_image(img)
    got = ctx.encode_level(dom, rng, fb.Params(0.0, -1.0, True))
    want = fo.encode_level(img, img, dom, rng, fo.params(0.0, -1.0, True))
    assert_items_equal(got, want)
    assert (got["distance"] == 100000.0).all() and (got["src_w"] == 0).all()


def test_quantizer_gpu(ctx, fo, lenna):
    z = np.load(os.path.join(GOLDEN, "items_lenna_16_8.npz"))["items"]
    qs, qo, mm = ctx.quantize(z, 5, 7)
    assert mm[0] == min(z["contrast"].min(), 1.7976931348623157e308) and mm[1] == max(z["contrast"].max(), -1.0)
    for i in range(0, len(z), 37):
        assert qs[i] == fo.quantize(z["contrast"][i], mm[0], mm[1], 5)
        assert qo[i] == fo.quantize(z["brightness"][i], mm[2], mm[3], 7)


def test_decode_mixed_sizes_and_stride(ctx, fo, lenna):
    """Quadtree list (mixed 16/8/4 items) decoded into a padded-stride plane, fixed iteration count."""
    z = np.load(os.path.join(GOLDEN, "items_lenna_qt_16_4_cls_thr5.npz"))["items"]
    for iters in (1, 3, 8):
        a, ia, ra = ctx.decode(z, 512, 512, stride=544, max_iters=iters, init=7)
        b, ib, rb = fo.decode(z, 512, 512, stride=544, max_iters=iters, init=7)
        assert (a == b).all() and ia == ib and ra == rb


def test_error_codes(ctx):
    import fractencode_b200 as fb
    img = np.zeros((32, 32), np.uint8)
    ctx.set_image(img)
    bad_dom = np.array([(0, 0, 8, 8, -1), (0, 0, 16, 16, -1)], fb.GRID_ITEM)
    rng = fb.uniform_grid(32, 32, 4, 4)
    with pytest.raises(fb.FractencodeError) as e:
        ctx.encode_level(bad_dom, rng, fb.Params())
    assert e.value.code == -2  # FE_ERR_UNSUPPORTED: mixed sizes
    out_dom = np.array([(28, 0, 8, 8, -1)], fb.GRID_ITEM)
    with pytest.raises(fb.FractencodeError) as e:
        ctx.encode_level(out_dom, rng, fb.Params())
    assert e.value.code == -1  # FE_ERR_INVALID: outside the image
    with pytest.raises(fb.FractencodeError):
        ctx.encode_quadtree(24, 4, fb.Params())  # not a power of two


def test_empty_and_degenerate_inputs(ctx, fo):
    """Empty range list, empty domain list (default items), a single range, image smaller than one domain."""
    import fractencode_b200 as fb
    img = fo.synth_image(64, 64, 3, 0)
    ctx.set_image(img)
    dom, rng = fb.uniform_grid(64, 64, 16, 8), fb.uniform_grid(64, 64, 8, 8)
    assert len(ctx.encode_level(dom, rng[:0], fb.Params())) == 0
    out = ctx.encode_level(dom[:0], rng, fb.Params(0.0, -1.0, True))
    assert (out["distance"] == 100000.0).all() and (out["src_w"] == 0).all() and (out["x"] == rng["x"]).all()
    one = ctx.encode_level(dom, rng[5:6], fb.Params())
    assert_items_equal(one, fo.encode_level(img, img, dom, rng[5:6], fo.params()))
    # quadtree on an image with no room for a 2T domain at the top level: the top level yields default items and splits
    small = fo.synth_image(32, 32, 3, 0)
    ctx.set_image(small)
    got, counts = ctx.encode_quadtree(32, 8, fb.Params(5.0))
    want, wcounts = fo.encode_quadtree(small, 32, 8, fo.params(5.0))
    assert counts == wcounts
    assert_items_equal(got, want)


def test_largest_block_size_and_padded_stride_quadtree(ctx, fo):
    """T = 64 (largest supported range block, exact integer path) and a quadtree on a plane with stride > width."""
    import fractencode_b200 as fb
    buf = np.zeros((256, 256 + 64), np.uint8)
    buf[:, :256] = fo.synth_image(256, 256, 11, 0)
    img = buf[:, :256]
    ctx.set_image(img)
    dom, rng = fb.uniform_grid(256, 256, 128, 64), fb.uniform_grid(256, 256, 64, 64)
    got = ctx.encode_level(dom, rng, fb.Params(0.0))
    assert_items_equal(got, fo.encode_level(img, img, dom, rng, fo.params(0.0)), "T=64")
    got, counts = ctx.encode_quadtree(32, 4, fb.Params(40.0, 1.5, True))
    want, wcounts = fo.encode_quadtree(img, 32, 4, fo.params(40.0, 1.5, True))
    assert counts == wcounts
    assert_items_equal(got, want, "padded-stride quadtree")
    with pytest.raises(fb.FractencodeError) as e:
        ctx.encode_level(fb.uniform_grid(256, 256, 256, 128), fb.uniform_grid(256, 256, 128, 128), fb.Params())
    assert e.value.code == -2  # T = 128 > 64


def test_packed_quantised_records_roundtrip(ctx, fo, lenna):
    """fe_pack_items / fe_unpack_items: q values and dequantised values equal the reference Quantizer's
    (oracle restatement of encode/Quantizer.hpp), geometry survives the 64-bit packing, and the unpacked list decodes."""
    z = np.load(os.path.join(GOLDEN, "items_lenna_qt_16_4_cls_thr5.npz"))["items"]
    packed, mm = ctx.pack_items(z, 16, 5, 7)
    assert packed.dtype == np.uint64 and len(packed) == len(z)
    assert mm[0] == z["contrast"].min() and mm[1] == max(z["contrast"].max(), -1.0)
    un = ctx.unpack_items(packed, 16, mm, 5, 7)
    for f in ("x", "y", "w", "h", "match_x", "match_y", "src_w", "src_h", "transform"):
        assert (un[f] == z[f]).all(), f
    qs, qo, mm2 = ctx.quantize(z, 5, 7)
    assert (mm2 == mm).all()
    assert (((packed >> np.uint64(49)) & np.uint64(31)) == qs).all() and (((packed >> np.uint64(54)) & np.uint64(127)) == qo).all()
    for i in range(0, len(z), 53):
        assert qs[i] == fo.quantize(z["contrast"][i], mm[0], mm[1], 5)
        assert un["contrast"][i] == fo.dequantize(int(qs[i]), mm[0], mm[1], 5)
        assert un["brightness"][i] == fo.dequantize(int(qo[i]), mm[2], mm[3], 7)
    # the quantised stream still decodes to something close to the image (5/7-bit coefficients)
    dec_q, _, _ = ctx.decode(un, 512, 512, max_iters=16)
    dec, _, _ = ctx.decode(z, 512, 512, max_iters=16)
    assert np.abs(dec_q.astype(int) - lenna.astype(int)).mean() < np.abs(dec.astype(int) - lenna.astype(int)).mean() + 4.0
    # items off the power-of-two lattice are refused
    bad = z[:4].copy()
    bad["x"] += 1
    import fractencode_b200 as fb
    with pytest.raises(fb.FractencodeError) as e:
        ctx.pack_items(bad, 16)
    assert e.value.code == -2
