import sys, time
sys.path.insert(0, '/root/repo')
import torch
import fractencode_b200 as fb
stream = torch.cuda.current_stream()
with fb.Context(0, stream.cuda_stream) as ctx:
    p = fb.Params(25.0, -1.0, False, False)
    for seed in range(1234, 1242):
        ctx.set_synthetic_image(4096, 4096, seed, 0)
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            n = ctx.encode_quadtree_slice_device(32, 4, p, 0, 16384)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        s = ctx.stats()
        print(seed, n, "ms %.3f" % min(ts[1:]), [int(s.level_evaluated[l]) for l in range(4)])
