#!/usr/bin/env python
"""Single-level search timing probe (one quadtree level T on a synthetic image), for kernel tuning / ncu."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractencode_b200 as fb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--T", type=int, default=4)
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--thr", type=float, default=-1.0)
ap.add_argument("--cls", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
with fb.Context(0) as ctx:
    ctx.set_synthetic_image(a.size, a.size, 1234, 0)
    p = fb.Params(a.thr, -1.0, bool(a.cls), False, a.impl)
    for rep in range(a.reps):
        ctx.stats_reset()
        n = ctx.encode_quadtree_device(a.T, a.T, p)
        st = ctx.stats()
        m, ms = st.level_matches[0], st.level_search_ms[0]
        print("T=%d size=%d impl=%d items=%d matches=%.3e search_ms=%.3f prep_ms=%.3f  %.1f TFLOP/s  %.2f Gmatch/s umma=%d dbg=%s" % (
            a.T, a.size, a.impl, n, m, ms, st.level_prep_ms[0], 2.0 * a.T * a.T * m / (ms * 1e-3) / 1e12, m / ms / 1e6, st.umma_levels,
            os.environ.get("FE_UMMA_DBG", "0")), flush=True)
