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


def _sampled_parity(fo, img, items, searched_by_T, W, H, thr, cls, per_level, seed):
    """Oracle check of `per_level` range blocks of EVERY searched level, each against the full domain grid of its level
    (reference rule: TransformEstimator2.hpp:29-48, Classifier2.cpp:70-81): a block the GPU emitted must carry the oracle's
    transform bit for bit; a block the GPU split must be one the oracle leaves above the threshold."""
    rs = np.random.default_rng(seed)
    emitted = {(int(e["x"]), int(e["y"]), int(e["w"])) for e in items[["x", "y", "w"]]}
    by_key = {(int(e["x"]), int(e["y"]), int(e["w"])): i for i, e in enumerate(items)}
    checked = {}
    for T, blocks in searched_by_T.items():
        pick = blocks[rs.choice(len(blocks), size=min(per_level, len(blocks)), replace=False)]
        dom = fo.uniform_grid(W, H, 2 * T, T)
        rng = np.zeros(len(pick), dom.dtype)
        rng["x"], rng["y"], rng["w"], rng["h"], rng["bin"] = pick[:, 0], pick[:, 1], T, T, -1
        if cls:
            dom, rng = fo.preclassify(img, dom), fo.preclassify(img, rng)
        want = fo.encode_level(img, img, dom, rng, fo.params(thr, -1.0, cls))
        kept = np.array([(int(x), int(y), T) in emitted for x, y in pick])
        can_split = T > 4
        if can_split:
            assert ((want["distance"] > thr) == ~kept).all(), "T=%d: split decisions differ from the oracle" % T
        else:
            assert kept.all()
        if kept.any():
            got = items[[by_key[(int(x), int(y), T)] for x, y in pick[kept]]]
            assert_items_equal(got, want[kept], "T=%d sample" % T)
        checked[T] = (int(kept.sum()), int((~kept).sum()))
    return checked


def _searched_blocks(items, W, H):
    """Range blocks searched per level, reconstructed from the emitted list: a block was searched at level T when it, or an
    emitted descendant of it, exists."""
    out = {}
    for T in (32, 16, 8, 4):
        sel = items[items["w"] <= T]
        keys = np.unique(np.stack([sel["x"] // T * T, sel["y"] // T * T], 1), axis=0)
        out[T] = keys.astype(np.int64)
    return out


@pytest.mark.parametrize("size,cls,thr,per_level", [(4096, False, 25.0, 256), (2048, True, 25.0, 64)])
def test_quadtree_full_size_properties_and_sampled_parity(ctx, fo, size, cls, thr, per_level):
    """BASELINE configs 3/4 in shape: 4096^2 quadtree 32->4 full search (the benchmark workload), and a classifier run."""
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
    blocks = _searched_blocks(items, W, H)
    assert [len(blocks[T]) for T in (32, 16, 8, 4)] == searched
    checked = _sampled_parity(fo, img, items, blocks, W, H, thr, cls, per_level, seed=5)
    assert all(k + s == min(per_level, len(blocks[T])) for T, (k, s) in checked.items())
    # decode round trip: the fixed-point iteration converges towards the image
    dec, it, rms = ctx.decode(items, W, H, max_iters=12, eps=-1e9)
    err = np.abs(dec.astype(np.int32) - img.astype(np.int32))
    assert err.mean() < 12.0, err.mean()


def test_config4_8192_classifier_sampled_parity_and_shard_union(ctx, fo):
    """BASELINE config 4 at its full size: 8192^2, quadtree 32->4, Classifier2 (classes x brightness bins, the default
    rule), sampled against the oracle on every level; and the union of 8 range shards is exactly the whole list."""
    import fractencode_b200 as fb
    W = H = 8192
    thr = 25.0
    ctx.set_synthetic_image(W, H, 4321, 0)
    img = ctx.get_image()
    p = fb.Params(thr, -1.0, True)
    n = ctx.encode_quadtree_device(32, 4, p)
    whole = ctx.fetch_items().copy()
    assert len(whole) == n
    _check_tiling(whole, W, H)
    blocks = _searched_blocks(whole, W, H)
    _sampled_parity(fo, img, whole, blocks, W, H, thr, True, 48, seed=11)
    n_top = (W // 32) * (H // 32)
    parts = []
    for r in range(8):
        lo, hi = r * n_top // 8, (r + 1) * n_top // 8
        ctx.encode_quadtree_slice_device(32, 4, p, lo, hi - lo)
        parts.append(ctx.fetch_items().copy())
    union = np.concatenate(parts)
    assert len(union) == len(whole)
    from oracle import pyoracle as po
    assert po.sort_items(union).tobytes() == po.sort_items(whole).tobytes()


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
