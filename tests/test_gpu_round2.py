"""Round-2 GPU tests: C-ABI hardening (trust-boundary checks), decode coverage proof and device-side convergence,
device-scheduled levels (slice hints + continuation), device-resident packed records."""
import os

import numpy as np
import pytest

from tests.cases import assert_items_equal
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_classifier_bins_out_of_range_are_rejected(ctx, fo):
    """Bins are -1 or Classifier2's classes 0..5; anything else is refused (it would index outside the bucket tables)."""
    import fractencode_b200 as fb
    img = fo.synth_image(64, 64, 5, 0)
    ctx.set_image(img)
    for bad in (6, 100, -2):
        for which in ("dom", "rng"):
            dom, rng = fb.uniform_grid(64, 64, 16, 8), fb.uniform_grid(64, 64, 8, 8)
            (dom if which == "dom" else rng)["bin"][3] = bad
            with pytest.raises(fb.FractencodeError) as e:
                ctx.encode_level(dom, rng, fb.Params(0.0, -1.0, True))
            assert e.value.code == -1, (bad, which)
            # without the classifier the field is ignored, as DummyClassifier does
            got = ctx.encode_level(dom, rng, fb.Params(0.0, -1.0, False))
            assert_items_equal(got, fo.encode_level(img, img, dom, rng, fo.params(0.0, -1.0, False)))


def test_bounds_checks_do_not_wrap_around(ctx, fo):
    """x + w is checked without 32-bit wrap-around at every entry point that takes block lists."""
    import fractencode_b200 as fb
    img = fo.synth_image(64, 64, 5, 0)
    ctx.set_image(img)
    dom, rng = fb.uniform_grid(64, 64, 16, 8), fb.uniform_grid(64, 64, 8, 8)
    for fld in ("x", "y"):
        r2 = rng.copy()
        r2[fld][1] = 0xFFFFFFFE
        with pytest.raises(fb.FractencodeError) as e:
            ctx.encode_level(dom, r2, fb.Params())
        assert e.value.code == -1
        d2 = dom.copy()
        d2[fld][1] = 0xFFFFFFF8
        with pytest.raises(fb.FractencodeError) as e:
            ctx.encode_level(d2, rng, fb.Params())
        assert e.value.code == -1
        with pytest.raises(fb.FractencodeError):
            ctx.classify(d2)
    items = ctx.encode_level(dom, rng, fb.Params())
    for fld in ("x", "y", "match_x", "match_y"):
        bad = items.copy()
        bad[fld][2] = 0xFFFFFFFC
        with pytest.raises(fb.FractencodeError) as e:
            ctx.decode(bad, 64, 64, max_iters=1)
        assert e.value.code == -1
        with pytest.raises(fb.FractencodeError):
            ctx.copy_items(img, img.copy(), bad)


def test_decode_equal_area_is_not_a_tiling(ctx, fo):
    """A list whose areas sum to the plane but which overlaps (and leaves a gap) must not take the ping-pong path."""
    import fractencode_b200 as fb
    img = fo.synth_image(64, 64, 9, 0)
    ctx.set_image(img)
    dom, rng = fb.uniform_grid(64, 64, 16, 8), fb.uniform_grid(64, 64, 8, 8)
    items = ctx.encode_level(dom, rng, fb.Params())
    dup = items.copy()
    dup[5] = dup[4]                               # item 4 twice, block 5 never written: same total area
    with pytest.raises(fb.FractencodeError) as e:
        ctx.decode(dup, 64, 64, max_iters=4)
    assert e.value.code == -1
    # a list with gaps only (no overlap) decodes like the reference: uncovered pixels keep the target's content
    gaps = np.delete(items, [3, 17, 40])
    for iters in (1, 2, 5, -1):
        a, ia, ra = ctx.decode(gaps, 64, 64, max_iters=iters, init=33)
        b, ib, rb = fo.decode(gaps, 64, 64, max_iters=iters, init=33)
        assert (a == b).all() and ia == ib and ra == rb, iters


def test_decode_leaves_the_callers_padding_alone(ctx, fo):
    import ctypes as C
    import fractencode_b200 as fb
    z = np.load(os.path.join(GOLDEN, "items_lenna_16_8.npz"))["items"]
    W = H = 512
    stride = 544
    tgt = np.full((H, stride), 7, np.uint8)
    it, rms = C.c_int(0), C.c_double(0)
    rc = ctx.lib.fe_decode(ctx.h, z.ctypes.data, len(z), tgt.ctypes.data, W, H, stride, 5, 1e-5, 0, C.byref(it), C.byref(rms))
    assert rc == 0
    assert (tgt[:, W:] == 7).all(), "padding bytes of the caller's plane were overwritten"
    ref, _, _ = fo.decode(z, W, H, stride=stride, max_iters=5, init=7)
    assert (tgt[:, :W] == ref).all()


def test_decode_converges_on_the_same_iteration_as_the_reference(ctx, fo, lenna):
    """Device-side convergence test in batches of eight iterations: iteration count, rms and image are the reference's
    whether the stop falls at the start, in the middle or at the end of a batch."""
    z = np.load(os.path.join(GOLDEN, "items_lenna_qt_16_4_cls_thr5.npz"))["items"]
    for eps in (1e-5, 0.5, 3.0, 40.0, 1e9):
        a, ia, ra = ctx.decode(z, 512, 512, eps=eps)
        b, ib, rb = fo.decode(z, 512, 512, eps=eps)
        assert ia == ib and ra == rb and (a == b).all(), (eps, ia, ib)
    for iters in (0, 1, 7, 8, 9, 16, 17):
        a, ia, ra = ctx.decode(z, 512, 512, max_iters=iters, eps=-1.0)
        b, ib, rb = fo.decode(z, 512, 512, max_iters=iters, eps=-1.0)
        assert ia == ib and ra == rb and (a == b).all(), iters


def test_decode_box_sum_plane_path_equals_plain_path_and_oracle(ctx, fo):
    """Lists of 4- and 8-pixel lattice blocks decode through the half-resolution box-sum plane (k_decode_step_small<.., DQ>);
    FE_NO_DQ=1 takes the pixel gather.  Same image, iteration count and rms either way and as the oracle -- with and without
    FMA, on a padded stride, for a list that also holds 16-pixel blocks (plane not used) and after a non-tiling call."""
    import fractencode_b200 as fb
    W, H = 384, 256
    ctx.set_synthetic_image(W, H, 4242, 0)
    img = ctx.get_image()
    for tmax in (8, 16):
        for thr in (25.0, 60.0, 120.0, 250.0):       # the first threshold that leaves both 8- and 4-pixel blocks
            items, counts = ctx.encode_quadtree(tmax, 4, fb.Params(thr))
            if counts[-1] > 0 and counts[-2] > 0:
                break
        assert counts[-1] > 0 and counts[-2] > 0, counts
        for fma in (False, True):
            for iters, eps in ((1, -1.0), (2, -1.0), (9, -1.0), (-1, 1e-5)):
                want, iw, rw = fo.decode(items, W, H, max_iters=iters, eps=eps, fma=fma)
                got = {}
                for name, env in (("dq", None), ("plain", "1")):
                    old = os.environ.pop("FE_NO_DQ", None)
                    if env:
                        os.environ["FE_NO_DQ"] = env
                    try:
                        got[name] = ctx.decode(items, W, H, max_iters=iters, eps=eps, fma=fma)
                    finally:
                        os.environ.pop("FE_NO_DQ", None)
                        if old is not None:
                            os.environ["FE_NO_DQ"] = old
                for name, (a, ia, ra) in got.items():
                    assert ia == iw and ra == rw and (a == want).all(), (tmax, fma, iters, name, ia, iw)
        a, ia, ra = ctx.decode(items, W, H, stride=W + 36, max_iters=3, eps=-1.0, init=9)
        b, ib, rb = fo.decode(items, W, H, stride=W + 36, max_iters=3, eps=-1.0, init=9)
        assert ia == ib and ra == rb and (a == b).all(), "padded stride"
    del img


def test_slice_hints_and_continuation(fo):
    """The number of slices a level gets enqueued up front comes from the previous level of its kind; a level that needs
    more is continued after its synchronisation.  Results never depend on what ran before on the context."""
    import fractencode_b200 as fb
    imgs = [fo.synth_image(256, 256, 1234, 2), fo.synth_image(256, 256, 1234, 0), fo.synth_image(256, 256, 7, 1),
            fo.synth_image(256, 256, 1234, 0)]
    params = [fb.Params(3.0), fb.Params(8.0), fb.Params(40.0, -1.0, True), fb.Params(8.0)]
    fresh = []
    for img, p in zip(imgs, params):
        with fb.Context(0) as c:
            c.set_image(img)
            fresh.append(c.encode_quadtree(32, 4, p))
    with fb.Context(0) as c:                      # one context: the hints of one image meet the next
        for rep in range(2):
            for (img, p), (want, wcounts) in zip(zip(imgs, params), fresh):
                c.set_image(img)
                got, counts = c.encode_quadtree(32, 4, p)
                assert counts == wcounts
                assert_items_equal(got, want)
    ref, rcounts = fo.encode_quadtree(imgs[1], 32, 4, fo.params(8.0))
    assert rcounts == fresh[1][1]
    assert_items_equal(fresh[1][0], ref)


def test_device_resident_pack_matches_the_host_pack(ctx, fo):
    """fe_items_minmax_device + fe_pack_items_device on the result list in HBM == fe_pack_items on the fetched list."""
    import torch
    import fractencode_b200 as fb
    img = fo.synth_image(256, 256, 1234, 0)
    ctx.set_image(img)
    n = ctx.encode_quadtree_device(32, 4, fb.Params(8.0))
    items = ctx.fetch_items()
    assert len(items) == n
    mm = torch.zeros(4, dtype=torch.float64, device="cuda")
    packed = torch.zeros(n + 5, dtype=torch.int64, device="cuda")
    ctx.items_minmax_device(mm.data_ptr())
    assert ctx.pack_items_device(32, mm.data_ptr(), packed.data_ptr(), n + 5) == n
    assert ctx.pack_errors() == 0
    want, wmm = ctx.pack_items(items, 32)
    assert (mm.cpu().numpy() == wmm).all()
    assert (packed.cpu().numpy()[:n].view(np.uint64) == want).all()
    # the decoder accepts what the packed stream carries
    back = ctx.unpack_items(want, 32, wmm)
    dec, it, rms = ctx.decode(back, 256, 256, max_iters=10, eps=-1.0)
    odec, oit, orms = fo.decode(back, 256, 256, max_iters=10, eps=-1.0)
    assert (dec == odec).all() and it == oit and rms == orms
    with pytest.raises(fb.FractencodeError):
        ctx.pack_items_device(32, mm.data_ptr(), packed.data_ptr(), n - 1)      # capacity


def test_batch_entry_point_small_images_vs_oracle(ctx, fo):
    """fe_encode_batch == the oracle's quadtree on every image (whole lists), incl. an odd batch size and a batch of one."""
    import fractencode_b200 as fb
    p = fb.Params(8.0, -1.0, True)
    for n in (1, 5):
        imgs = [fo.synth_image(256, 256, 100 + i, i % 2) for i in range(n)]
        lists = ctx.encode_batch(imgs, 32, 4, p)
        assert len(lists) == n
        for img, got in zip(imgs, lists):
            want, _ = fo.encode_quadtree(img, 32, 4, fo.params(8.0, -1.0, True))
            assert_items_equal(got, want)


def test_batch_of_16_images_1024(ctx, fo):
    """BASELINE config 5 in shape: sixteen 1024^2 images, 8x8 grid + quadtree split to 4x4, through fe_encode_batch.
    Every list equals the single-image call; image 0 is compared WHOLE with the oracle, the others on a sample."""
    import fractencode_b200 as fb
    from oracle import pyoracle as po
    thr = 25.0
    p = fb.Params(thr)
    imgs = [fo.synth_image(1024, 1024, 1234 + i, 0) for i in range(16)]
    lists = ctx.encode_batch(imgs, 8, 4, p)
    assert len(lists) == 16
    rs = np.random.default_rng(3)
    for i, (img, got) in enumerate(zip(imgs, lists)):
        ctx.set_image(img)
        single, _ = ctx.encode_quadtree(8, 4, p)
        assert po.sort_items(got).tobytes() == po.sort_items(single).tobytes(), i
        if i == 0:
            want, _ = fo.encode_quadtree(img, 8, 4, fo.params(thr))
            assert_items_equal(got, want, "image 0 whole")
        else:
            for T in (8, 4):
                sel = got[got["w"] == T]
                pick = sel[rs.choice(len(sel), size=min(24, len(sel)), replace=False)]
                dom = fo.uniform_grid(1024, 1024, 2 * T, T)
                rng = np.zeros(len(pick), dom.dtype)
                rng["x"], rng["y"], rng["w"], rng["h"], rng["bin"] = pick["x"], pick["y"], T, T, -1
                assert_items_equal(pick, fo.encode_level(img, img, dom, rng, fo.params(thr)), "image %d T=%d" % (i, T))


@pytest.mark.parametrize("kind,S,T,thr,cls", [(0, 16, 8, 0.0, False), (0, 8, 4, 0.0, True), (0, 32, 16, 0.0, False), (2, 16, 8, 0.0, False),
                                               (2, 16, 8, 3.0, True), (0, 16, 8, 30.0, False), (0, 32, 16, 60.0, True)])
def test_flip_isometries_fixed_grid(ctx, fo, kind, S, T, thr, cls):
    """the match chain extended through the four flip isometries (image/transform.h:20-24,37-40), behind
    fe_params.isometries = 8.  Oracle = the same chain rules over eight isometries (oracle/frac_oracle.c:match_chain)."""
    import fractencode_b200 as fb
    img = fo.synth_image(128, 128, 77, kind)
    dom, rng = fb.uniform_grid(128, 128, S, S // 2), fb.uniform_grid(128, 128, T, T)
    ctx.set_image(img)
    for fma in (False, True):
        got = ctx.encode_level(dom, rng, fb.Params(thr, -1.0, cls, fma, isometries=8))
        want = fo.encode_level(img, img, dom, rng, fo.params(thr, -1.0, cls, fma, isometries=8))
        assert_items_equal(got, want, "flips S=%d T=%d" % (S, T))
    if kind != 2:   # (the pattern image matches exactly under the identity)
        assert (got["transform"] > 3).any(), "no flip isometry was ever selected: the test would prove nothing"
    # and the reference's four-rotation answer is untouched by the extension
    got4 = ctx.encode_level(dom, rng, fb.Params(thr, -1.0, cls))
    assert_items_equal(got4, fo.encode_level(img, img, dom, rng, fo.params(thr, -1.0, cls)))


def test_flip_isometries_quadtree_and_decode(ctx, fo):
    import fractencode_b200 as fb
    img = fo.synth_image(256, 256, 1234, 0)
    ctx.set_image(img)
    for cls in (False, True):
        got, counts = ctx.encode_quadtree(32, 4, fb.Params(8.0, -1.0, cls, isometries=8))
        want, wcounts = fo.encode_quadtree(img, 32, 4, fo.params(8.0, -1.0, cls, isometries=8))
        assert counts == wcounts
        assert_items_equal(got, want)
        dec, it, rms = ctx.decode(got, 256, 256)
        odec, oit, orms = fo.decode(want, 256, 256)
        assert (dec == odec).all() and it == oit and rms == orms
    # more isometries can only find hits earlier: never more items than with rotations alone... not guaranteed by the first-hit
    # rule, but the candidate count doubles
    ctx.stats_reset()
    ctx.encode_quadtree(32, 4, fb.Params(8.0, isometries=8))
    m8 = int(ctx.stats().level_matches[0])
    ctx.stats_reset()
    ctx.encode_quadtree(32, 4, fb.Params(8.0))
    assert m8 == 2 * int(ctx.stats().level_matches[0])
    # the fp32-rounding regime (noise-like blocks at T >= 16) is re-ranked over the four rotations only: refused with flips
    noise = fo.synth_image(128, 128, 77, 1)
    ctx.set_image(noise)
    with pytest.raises(fb.FractencodeError) as e:
        ctx.encode_level(fb.uniform_grid(128, 128, 32, 16), fb.uniform_grid(128, 128, 16, 16), fb.Params(0.0, isometries=8))
    assert e.value.code == -2
    ctx.set_image(img)
    # generic geometries (exact integer path) do not take the flag
    dom, rng = fb.uniform_grid(256, 256, 16, 8), fb.uniform_grid(256, 256, 4, 4)
    with pytest.raises(fb.FractencodeError) as e:
        ctx.encode_level(dom, rng, fb.Params(0.0, isometries=8))
    assert e.value.code == -2
