#!/usr/bin/env python
"""Randomised A/B of the pruned threshold search against the plain one-pass search (FE_SINGLE_PASS=1) and, on small images,
against the exact integer (dp4a) kernel, bit for bit.
usage: stress_pruning.py [seconds] [seed] [max edge / 64]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractencode_b200 as fb  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (sort_items only)

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rs = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
max_blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 16
FIELDS = ("x", "y", "w", "h", "match_x", "match_y", "src_w", "src_h", "transform")
t0 = time.time()
runs = 0
with fb.Context(0) as ctx:
    while time.time() - t0 < budget:
        W = int(rs.integers(2, max_blocks + 1)) * 64
        H = int(rs.integers(2, max_blocks + 1)) * 64
        kind = int(rs.integers(0, 3))
        tmax = int(rs.choice([8, 16, 32, 64]))
        tmin = int(rs.choice([t for t in (4, 8, 16, 32) if t <= tmax]))
        thr = float(rs.choice([0.0, 0.5, 3.0, 10.0, 25.0, 60.0, 150.0, 400.0]) * rs.uniform(0.5, 1.5))
        cls = bool(rs.integers(0, 2))
        seed = int(rs.integers(0, 1 << 30))
        single = bool(rs.integers(0, 4) == 0)       # sometimes a single level through fe_encode_level (minimum always wanted)
        ctx.set_synthetic_image(W, H, seed, kind)
        if rs.integers(0, 3) == 0:                  # same pixels through a host plane with an odd stride (unaligned rows)
            img = ctx.get_image()
            pad = np.zeros((H, W + int(rs.integers(1, 9))), np.uint8)
            pad[:, :W] = img
            ctx.set_image(pad[:, :W])
        p = fb.Params(thr, -1.0, cls)
        out = {}
        modes = ("pruned", "one_pass", "exact") if W * H <= 512 * 512 else ("pruned", "one_pass")   # exact = dp4a integer kernel
        try:
            for mode in modes:
                p = fb.Params(thr, -1.0, cls, False, 1 if mode == "exact" else 0)
                if mode == "one_pass":
                    os.environ["FE_SINGLE_PASS"] = "1"
                else:
                    os.environ.pop("FE_SINGLE_PASS", None)
                if single:
                    dom, rng = fb.uniform_grid(W, H, 2 * tmin, tmin), fb.uniform_grid(W, H, tmin, tmin)
                    rng = rng[:: int(rs.integers(1, 4))] if mode == "pruned" else out["rng"]
                    out["rng"] = rng
                    out[mode] = ctx.encode_level(dom, rng, p)
                else:
                    out[mode] = ctx.encode_quadtree(tmax, tmin, p)[0]
        except fb.FractencodeError as e:
            if "fp32-rounding regime" in str(e) or "not aligned" in str(e):
                continue
            print("ERROR", dict(W=W, H=H, kind=kind, tmax=tmax, tmin=tmin, thr=thr, cls=cls, seed=seed, single=single, mode=mode), e, flush=True)
            sys.exit(2)
        finally:
            os.environ.pop("FE_SINGLE_PASS", None)
        ok = True
        for other in modes[1:]:
            a, b = po.sort_items(out["pruned"]), po.sort_items(out[other])
            ok = ok and len(a) == len(b) and all((a[f] == b[f]).all() for f in FIELDS) and all(
                (a[f].view(np.uint64) == b[f].view(np.uint64)).all() for f in ("distance", "contrast", "brightness"))
        runs += 1
        if not ok:
            print("MISMATCH", dict(W=W, H=H, kind=kind, tmax=tmax, tmin=tmin, thr=thr, cls=cls, seed=seed, single=single), flush=True)
            sys.exit(1)
print("stress_pruning: %d random configurations identical in %.0f s" % (runs, time.time() - t0))
