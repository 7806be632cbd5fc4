#!/usr/bin/env python
"""Short-item experiment: one full-scan launch of a T = 8 level (1024^2: 512 row tiles x 127 column tiles), work items cut to
FE_F16_MIN_RUN column tiles (FE_F16_ITEMS work items per SM).  Prints the kernel time.  (With a build whose A builders only
signed their barriers after the first tile of each buffer -- stale but real rows -- the same runs were 6-8 % faster at every
run length from 21 to 87 tiles -- proportional to time rather than to the number of items, so probably the epilogue's
data-dependent paths on stale rows rather than the cost of the build.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractencode_b200 as fb  # noqa: E402
os.environ["FE_SINGLE_PASS"] = "1"
with fb.Context(0) as ctx:
    ctx.set_synthetic_image(2048, 1024, 99, 0)
    for rep in range(3):
        ctx.stats_reset()
        ctx.encode_quadtree(8, 8, fb.Params(25.0))
        s = ctx.stats()
    print({k: os.environ.get(k) for k in ("FE_F16_ITEMS", "FE_F16_MIN_RUN", "FE_NO_PAIR")}, "search_ms %.4f" % s.level_search_ms[0],
          "evaluated %.3e" % s.level_evaluated[0], "G/ms %.2f" % (s.level_evaluated[0] / s.level_search_ms[0] / 1e9))
