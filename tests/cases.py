"""Shared helpers: run a golden case (tests/golden/goldens.json) on an oracle library or on the CUDA path."""
import hashlib

import numpy as np

from oracle import pyoracle as po


def md5_dump(items) -> str:
    return hashlib.md5(po.dump_lines(items).encode()).hexdigest()


def run_case_oracle(lib, img, c, fma=False, nthreads=0):
    H, W = img.shape
    p = lib.params(c.get("thr", 0.0), c.get("smax", -1.0), c.get("cls", False), fma)
    if c.get("qt"):
        return lib.encode_quadtree(img, c["qt"][0], c["qt"][1], p, nthreads)
    dom = lib.uniform_grid(W, H, c["S"], c["S"] // 2)
    rng = lib.uniform_grid(W, H, c["T"], c["T"])
    if c.get("cls", False):
        dom, rng = lib.preclassify(img, dom), lib.preclassify(img, rng)
    return lib.encode_level(img, img, dom, rng, p, nthreads), None


def run_case_gpu(ctx, img, c, fma=False, search_impl=0):
    import fractencode_b200 as fb
    H, W = img.shape
    p = fb.Params(c.get("thr", 0.0), c.get("smax", -1.0), c.get("cls", False), fma, search_impl)
    ctx.set_image(img)
    if c.get("qt"):
        return ctx.encode_quadtree(c["qt"][0], c["qt"][1], p)
    dom = fb.uniform_grid(W, H, c["S"], c["S"] // 2)
    rng = fb.uniform_grid(W, H, c["T"], c["T"])
    return ctx.encode_level(dom, rng, p), None


def assert_items_equal(got, want, what=""):
    """Bit-exact comparison of two transform lists (sorted by (y,x,w,h))."""
    got, want = po.sort_items(np.asarray(got)), po.sort_items(np.asarray(want))
    assert len(got) == len(want), "%s: %d items vs %d" % (what, len(got), len(want))
    for f in ("x", "y", "w", "h", "match_x", "match_y", "src_w", "src_h", "transform"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, "%s: field %s differs at %d items, first %s: got %s want %s" % (
            what, f, bad.size, bad[:3], got[bad[:3]], want[bad[:3]])
    for f in ("distance", "contrast", "brightness"):
        bad = np.nonzero(got[f].view(np.uint64) != want[f].view(np.uint64))[0]
        assert bad.size == 0, "%s: %s bits differ at %d items, first: got %r want %r" % (
            what, f, bad.size, got[bad[:3]], want[bad[:3]])
