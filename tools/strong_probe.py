#!/usr/bin/env python
"""Strong-scaling probe on ONE GPU: encode shard 0 of `--world` of the config-4 image and print the per-level split, next to the
whole image.  Shows which part of a rank's time does not shrink with the shard."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import fractencode_b200 as fb  # noqa: E402
from fractencode_b200.dist import shard_slice  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--classifier", type=int, default=1)
a = ap.parse_args()
stream = torch.cuda.current_stream()
with fb.Context(0, stream.cuda_stream) as ctx:
    ctx.set_synthetic_image(a.size, a.size, 4321, 0)
    p = fb.Params(25.0, -1.0, bool(a.classifier), False)
    n_top = (a.size // 32) ** 2
    for world in (1, a.world):
        sl = shard_slice(n_top, 0, world)
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = ctx.encode_quadtree_slice_device(32, 4, p, sl.start, sl.stop - sl.start)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
        s = ctx.stats()
        print("world=%d blocks=%d items=%d wall_ms=%.3f" % (world, sl.stop - sl.start, n, ms))
        for l in range(4):
            print("   T=%2d ranges=%8d evaluated=%14d passes=%2d search_ms=%.3f prep_ms=%.3f" % (32 >> l, s.level_ranges[l], s.level_evaluated[l], s.level_passes[l],
                                                                                                  s.level_search_ms[l], s.level_prep_ms[l]))
