#!/usr/bin/env python
"""Small end-to-end run touching every kernel (for compute-sanitizer): quadtree + classifier + decode + quantize + generic geometry."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractencode_b200 as fb  # noqa: E402

with fb.Context(0) as ctx:
    ctx.set_synthetic_image(256, 192, 7, 0)
    for cls in (False, True):
        items, counts = ctx.encode_quadtree(32, 4, fb.Params(30.0, -1.0, cls))
        print("quadtree cls=%d" % cls, counts, len(items))
    dec, it, rms = ctx.decode(items, 256, 192, max_iters=5)
    qs, qo, mm = ctx.quantize(items)
    dom, rng = fb.uniform_grid(256, 192, 16, 8), fb.uniform_grid(256, 192, 4, 4)   # rho = 4: exact path, 4 pools
    out = ctx.encode_level(dom, rng[:200], fb.Params(0.0))
    ctx.set_synthetic_image(128, 128, 7, 1)                                        # noise: fp32-regime re-rank at T=32
    out2 = ctx.encode_level(fb.uniform_grid(128, 128, 64, 32), fb.uniform_grid(128, 128, 32, 32), fb.Params(0.0))
    print("ok", it, mm, out["distance"][:2], out2["distance"][:2], ctx.stats().fp32_regime_items)
