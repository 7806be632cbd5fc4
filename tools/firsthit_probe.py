#!/usr/bin/env python
"""Where in the domain scan does the first candidate under the threshold sit?  (sizing the multi-pass early-out)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractencode_b200 as fb  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 25.0
with fb.Context(0) as ctx:
    ctx.set_synthetic_image(size, size, 1234, 0)
    p = fb.Params(thr, -1.0, False, False, 0)
    items, counts = ctx.encode_quadtree(32, 4, p)
    for T in (32, 16, 8, 4):
        it = items[items["w"] == T]
        if not len(it):
            print("T=%d: no items" % T)
            continue
        S = 2 * T
        nx = (size - S) // T + 1
        nD = nx * nx
        hit = it[it["distance"] <= thr]
        d = (hit["match_y"] // T).astype(np.int64) * nx + hit["match_x"] // T
        q = np.quantile(d / nD, [0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 1.0]) if len(d) else []
        print("T=%d items=%d hits=%d (%.1f%%) nD=%d first-hit position quantiles (fraction of scan) 10/25/50/75/90/99/100%%: %s  mean=%.4f"
              % (T, len(it), len(hit), 100.0 * len(hit) / len(it), nD, np.round(q, 4), (d / nD).mean() if len(d) else -1))
        # tile-level view: 32 consecutive ranges in item order share a row tile
        for frac in (1 / 64, 1 / 16, 1 / 4):
            print("    hits with d < %.4f of scan: %.1f%% of level items" % (frac, 100.0 * np.count_nonzero(d < frac * nD) / len(it)))
