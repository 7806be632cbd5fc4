"""BASELINE-size runs (the oracle cannot finish them): size-independent properties plus an oracle check of a
deterministic sample of the emitted items against the FULL domain grid of their level."""
import numpy as np
import pytest

from tests.cases import assert_items_equal

pytestmark = pytest.mark.gpu


def _check_tiling(items, W, H):
    cover = np.zeros((H, W), np.uint8)
    for T in np.unique(items["w"]):
        sel = items[items["w"] == T]
        assert (sel["h"] == T).all() and (sel["x"] % T == 0).all() and (sel["y"] % T == 0).all()
        idx = np.zeros((H // T, W // T), np.uint8)
        np.add.at(idx, (sel["y"] // T, sel["x"] // T), 1)
        cover += np.kron(idx, np.ones((T, T), np.uint8))
    assert (cover == 1).all(), "items must tile the plane exactly once"


@pytest.mark.parametrize("size,cls,thr", [(4096, False, 25.0), (2048, True, 25.0)])
def test_quadtree_full_size_properties_and_sampled_parity(ctx, fo, size, cls, thr):
    """BASELINE configs 3/4 in shape: 4096^2 quadtree 32->4 full search, and a classifier run."""
    import fractencode_b200 as fb
    W = H = size
    ctx.set_synthetic_image(W, H, 1234, 0)
    img = ctx.get_image()
    assert (img[:64, :64] == fo.synth_image(64, 64, 1234, 0)).all()
    p = fb.Params(thr, -1.0, cls)
    ctx.stats_reset()
    items, counts = ctx.encode_quadtree(32, 4, p)
    st = ctx.stats()
    assert st.exact_levels == 0 and st.umma_levels == 4
    assert sum(counts) == len(items)
    _check_tiling(items, W, H)
    # threshold semantics: an item above the threshold can only sit on the last level; split decisions are consistent
    assert (items["distance"][items["w"] > 4] <= thr).all()
    assert (items["src_w"] == 2 * items["w"]).all() and (items["transform"] >= 0).all() and (items["transform"] <= 3).all()
    # candidate count: ranges searched per level = 4 x (blocks that split on the level above)
    searched = [int(st.level_ranges[l]) for l in range(4)]
    assert searched[0] == (W // 32) * (H // 32)
    for l in range(1, 4):
        assert searched[l] == 4 * (searched[l - 1] - counts[l - 1])
    if not cls:
        nd = [(W // T - 1) * (H // T - 1) for T in (32, 16, 8, 4)]
        assert [int(st.level_matches[l]) for l in range(4)] == [searched[l] * nd[l] * 4 for l in range(4)]
    # sampled parity against the oracle: a few emitted items per level, each against the full domain grid of its level
    rs = np.random.default_rng(5)
    for T in np.unique(items["w"]):
        sel = items[items["w"] == T]
        pick = sel[rs.choice(len(sel), size=min(6, len(sel)), replace=False)]
        dom = fo.uniform_grid(W, H, 2 * int(T), int(T))
        rng = np.zeros(len(pick), dom.dtype)
        rng["x"], rng["y"], rng["w"], rng["h"], rng["bin"] = pick["x"], pick["y"], T, T, -1
        if cls:
            dom = fo.preclassify(img, dom)
            rng = fo.preclassify(img, rng)
        want = fo.encode_level(img, img, dom, rng, fo.params(thr, -1.0, cls))
        assert_items_equal(pick, want, "T=%d sample" % T)
    # decode round trip: the fixed-point iteration converges towards the image
    dec, it, rms = ctx.decode(items, W, H, max_iters=12, eps=-1e9)
    err = np.abs(dec.astype(np.int32) - img.astype(np.int32))
    assert err.mean() < 12.0, err.mean()


def test_batch_mode_images_are_independent(ctx, fo):
    """BASELINE config 5 in miniature: a batch of 1024^2 images, 8x8 grid; image i encodes the same alone or in a batch."""
    import fractencode_b200 as fb
    dom, rng = fb.uniform_grid(1024, 1024, 16, 8), fb.uniform_grid(1024, 1024, 8, 8)[::97]
    outs = []
    for seed in (1234, 1235, 1234):
        ctx.set_synthetic_image(1024, 1024, seed, 0)
        outs.append(ctx.encode_level(dom, rng, fb.Params(0.0)))
    assert outs[0].tobytes() == outs[2].tobytes() and outs[0].tobytes() != outs[1].tobytes()
    img = fo.synth_image(1024, 1024, 1234, 0)
    want = fo.encode_level(img, img, fo.uniform_grid(1024, 1024, 16, 8), rng[:8], fo.params(0.0))
    assert_items_equal(outs[0][:8], want)
