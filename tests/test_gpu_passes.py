"""Threshold searches prune the scan (slices of the domain order with early-out, brightness bins, hit-only bookkeeping on
levels that split).  None of it may change a single bit of the result: the pruned search must equal the plain one-pass
search (FE_SINGLE_PASS=1, every range against every admissible domain) and the oracle."""
import os

import numpy as np
import pytest

from tests.cases import assert_items_equal

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# default = slices + brightness bins (+ the lower-bound prefilter on the T = 32 level); "no_prefilter" takes the exact kind there;
# "one_pass" = one full scan per level, no pruning of any kind
# "plain_prep" = one row tile per work item and per-block class / bin / norm kernels instead of the level's cell sums
MODES = {"pruned": {}, "no_prefilter": {"FE_NO_LB": "1"}, "slices_only": {"FE_NO_BINS": "1", "FE_NO_LB": "1"},
         "plain_prep": {"FE_NO_PAIR": "1", "FE_NO_CELLS": "1"}, "one_pass": {"FE_SINGLE_PASS": "1", "FE_NO_LB": "1"}}


@pytest.mark.parametrize("kind,cls,thr", [(0, False, 25.0), (0, False, 6.0), (0, True, 25.0), (1, False, 40.0), (2, False, 25.0)])
def test_quadtree_pruned_equals_one_pass(ctx, kind, cls, thr):
    """1024^2, quadtree 32->4: natural / noise / pattern images, with and without classifier."""
    import fractencode_b200 as fb
    W = H = 1024
    ctx.set_synthetic_image(W, H, 77, kind)
    out = {}
    for name, env in MODES.items():
        with _Env(**env):
            ctx.stats_reset()
            items, counts = ctx.encode_quadtree(32, 4, fb.Params(thr, -1.0, cls))
            st = ctx.stats()
            out[name] = (items, counts, int(st.matches), int(st.evaluated))
    ref_items, ref_counts, ref_matches, ref_eval = out["one_pass"]
    assert ref_eval == ref_matches, "the one-pass search scores every admissible candidate"
    for name in ("pruned", "no_prefilter", "slices_only", "plain_prep"):
        items, counts, matches, evaluated = out[name]
        assert counts == ref_counts and matches == ref_matches
        # worst case on the last level: the bins (about a third of the scan) found hits for some ranges only, the rest
        # still needs the plain pass for its minimum
        assert evaluated <= 1.35 * matches
        assert_items_equal(items, ref_items, name)


@pytest.mark.parametrize("T,thr", [(4, 12.0), (8, 30.0), (16, 60.0), (32, 90.0)])
def test_single_level_with_threshold_keeps_the_minimum(ctx, fo, T, thr):
    """fe_encode_level cannot split: ranges without a hit need the minimum over ALL domains (the plain pass after the bins)."""
    import fractencode_b200 as fb
    W, H = 1024, 512
    ctx.set_synthetic_image(W, H, 5, 0)
    img = ctx.get_image()
    dom, rng = fb.uniform_grid(W, H, 2 * T, T), fb.uniform_grid(W, H, T, T)
    if T == 4:
        rng = rng[::5]
    got = {}
    for name, env in MODES.items():
        with _Env(**env):
            ctx.stats_reset()
            got[name] = ctx.encode_level(dom, rng, fb.Params(thr))
            st = ctx.stats()
            assert st.umma_levels == 1
            if name == "one_pass":
                assert int(st.evaluated) == int(st.matches)
    assert_items_equal(got["pruned"], got["one_pass"], "T=%d pruned" % T)
    assert_items_equal(got["slices_only"], got["one_pass"], "T=%d slices" % T)
    assert_items_equal(got["plain_prep"], got["one_pass"], "T=%d plain prep" % T)
    hits = np.count_nonzero(got["one_pass"]["distance"] <= thr)
    assert hits > 0 and (T == 32 or hits < len(rng)), "the case must mix ranges with and without a hit (%d of %d)" % (hits, len(rng))
    sub = rng[:: max(1, len(rng) // 16)][:16]
    want = fo.encode_level(img, img, dom, sub, fo.params(thr))
    assert_items_equal(ctx.encode_level(dom, sub, fb.Params(thr)), want, "T=%d oracle sample" % T)


def test_first_hit_is_first_in_scan_order_across_bins(ctx, fo):
    """Two flat halves one grey level either side of a bin border: every domain is under the threshold for every range, and
    the ranges of the right half (the lower bin) meet their own bin's domains BEFORE domain 0 in the operand layout.  The
    winner must still be domain 0 / rotation 0 -- the first of the scan -- not the first of some bin."""
    import fractencode_b200 as fb
    W = H = 512
    img = np.full((H, W), 101, np.uint8)   # sum(4r) = 404 N: bin 5 of width 80 N + 1 at rms_threshold 100
    img[:, W // 2:] = 99                   # 396 N: bin 4
    ctx.set_image(img)
    T = 8
    dom, rng = fb.uniform_grid(W, H, 2 * T, T), fb.uniform_grid(W, H, T, T)
    got = ctx.encode_level(dom, rng, fb.Params(100.0))
    want = fo.encode_level(img, img, dom, rng, fo.params(100.0))
    assert_items_equal(got, want, "flat image")
    assert (got["match_x"] == 0).all() and (got["match_y"] == 0).all() and (got["transform"] == 0).all()


@pytest.mark.parametrize("cls", [False, True])
def test_quadtree_slices_give_the_whole_transform_list(ctx, cls):
    """Range sharding: the shards of a partition of the top-level grid, each searched against the domains of the
    whole image, together are exactly the transform list of the whole image."""
    import fractencode_b200 as fb
    from fractencode_b200.dist import shard_slice
    W, H = 1024, 512
    ctx.set_synthetic_image(W, H, 99, 0)
    p = fb.Params(20.0, -1.0, cls)
    whole, counts = ctx.encode_quadtree(32, 4, p)
    n_top = (W // 32) * (H // 32)
    parts = []
    for r in range(3):
        sl = shard_slice(n_top, r, 3)
        n = ctx.encode_quadtree_slice_device(32, 4, p, sl.start, sl.stop - sl.start)
        got = ctx.fetch_items()
        assert len(got) == n
        parts.append(got.copy())
    assert sum(len(x) for x in parts) == len(whole)
    assert_items_equal(np.concatenate(parts), whole, "union of three shards")
    assert ctx.encode_quadtree_slice_device(32, 4, p, n_top, 0) == 0
    with pytest.raises(fb.FractencodeError):
        ctx.encode_quadtree_slice_device(32, 4, p, n_top - 1, 2)


def test_short_work_items_many_small_buckets(ctx):
    """Classifier buckets with fewer than four column tiles: work items in which some MMA issuers have no tile at all (a
    release/reuse race of the A-tile buffers showed up here as winners that failed the self-check, found by
    tools/stress_pruning.py).  Repeated, because a race does not fail every time."""
    import fractencode_b200 as fb
    W, H = 512, 768
    ctx.set_synthetic_image(W, H, 434051177, 0)
    p = fb.Params(31.10872017449507, -1.0, True)
    with _Env(FE_SINGLE_PASS="1"):
        want, _ = ctx.encode_quadtree(32, 8, p)
    for rep in range(12):
        got, _ = ctx.encode_quadtree(32, 8, p)
        assert_items_equal(got, want, "repetition %d" % rep)
