#!/usr/bin/env python
"""Decode timing probe: quadtree-encode a synthetic image, then run a fixed number of decode iterations."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractencode_b200 as fb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--thr", type=float, default=25.0)
ap.add_argument("--tmax", type=int, default=32)
ap.add_argument("--tmin", type=int, default=4)
ap.add_argument("--iters", type=int, default=8)
a = ap.parse_args()
with fb.Context(0) as ctx:
    ctx.set_synthetic_image(a.size, a.size, 1234, 0)
    items, counts = ctx.encode_quadtree(a.tmax, a.tmin, fb.Params(a.thr))
    for rep in range(3):
        img, it, rms = ctx.decode(items, a.size, a.size, max_iters=a.iters, eps=-1e9)
        ms = ctx.stats().last_decode_ms
        nbytes = 2.0 * a.size * a.size + 64.0 * len(items)
        print("items=%d %s iters=%d decode_ms=%.3f per_iter_ms=%.4f  %.0f GB/s algorithmic" % (len(items), counts, it, ms, ms / a.iters, nbytes / (ms / a.iters * 1e-3) / 1e9), flush=True)
