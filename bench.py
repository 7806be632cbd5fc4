#!/usr/bin/env python
"""bench.py -- fractal-encoding search throughput on N B200s of one node, or the reference CPU arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path over one synthetic image per GPU: the whole quadtree encode
(BASELINE config 3: 4096x4096 "natural" image, quadtree 32->4, full search) -- operand preparation,
the (range, domain x rotation) search on every level, winner finalisation with the least-squares
(s, o), and the split/compaction between levels.  `value` times it with the image already in HBM and
the transform list left in HBM (+ the NCCL gather of the per-rank lists, as packed 8-byte records, to rank 0 for N > 1); `e2e` times the
reference-facing C-ABI call with HOST buffers (pinned image in, transform list out).

matches = sum over levels of (range blocks searched) x (domains) x 4 rotations: the FULL candidate
count of the workload, identical for both arms, so matches/s ratios are time ratios.  With a
threshold both arms stop a range block at its first candidate under it (the reference's `break`,
TransformEstimator2.hpp:40-41; here at pass granularity), so fewer candidates are actually scored:
`evaluated` counts those, and every roofline figure is computed from `evaluated`, never from `matches`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def measured_traffic(a, T):
    """dram__bytes_read.sum + dram__bytes_write.sum of the level's search launches from the committed ncu --set full capture of
    exactly this workload (profiles/traffic_r2.json, written by tools/traffic_from_ncu.py); (None, reason) when there is none."""
    path = os.path.join(ROOT, "profiles", "traffic_r2.json")
    key = "%d/%d/%d/%g/%d" % (a.size, a.tmax, a.tmin, a.thr, a.classifier)
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d[key]["dram_bytes_by_T"][str(T)]), d[key].get("source", "profiles/traffic_r2.json")
    except Exception:
        return None, "no ncu capture committed for this workload"


METRIC = "range_block_matches_per_s"
UNIT = "matches/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--tmax", type=int, default=32)
    ap.add_argument("--tmin", type=int, default=4)
    ap.add_argument("--thr", type=float, default=float(os.environ.get("FE_BENCH_THR", "25")))
    ap.add_argument("--classifier", type=int, default=0)
    ap.add_argument("--search", type=int, default=0, help="0 auto, 1 exact integer path, 2 tcgen05 path")
    ap.add_argument("--cpu-blocks", type=int, default=0, help="range blocks per level in the CPU sample (0: 2 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch", action="store_true", help="skip the batch-mode sub-measurement (config 5)")
    ap.add_argument("--no-decode-large", action="store_true", help="N = 1: skip the 8192 x 8192 decode measurement")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling sub-measurement (config 4)")
    ap.add_argument("--shard", default="images", choices=["images", "ranges"],
                    help="N > 1: one image per GPU (weak scaling, default) or the range blocks of ONE image sharded over the GPUs "
                         "(strong scaling, BASELINE config 4 in shape; the domain pool is rebuilt on every GPU)")
    return ap.parse_args()


def workload_name(a):
    return "natural %dx%d u8 (SURVEY 8d generator, seed 1234+rank), quadtree %d->%d, %s, rms_threshold %g" % (
        a.size, a.size, a.tmax, a.tmin, "Classifier2" if a.classifier else "full search (DummyClassifier)", a.thr)


PRUNING_NOTE = ("exact: range blocks stop at their first candidate under the threshold (slices of the scan, reference break semantics) and, "
                "without classifier, meet only domains whose pixel sum can satisfy the threshold (Cauchy-Schwarz bins); "
                "`matches` is the nominal candidate count, `evaluated` what the kernels scored")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "tflops_burst": d.get("bf16_tflops", 1590.0),
                "tflops_sustained": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the real reference (oracle/_ref) when it was compiled, else the oracle port.
# Hierarchical bounded sample: m blocks of the top level against the FULL domain grid; the children of
# the sampled blocks that split are the sample of the next level; per-level cost and split fraction are
# scaled to the whole image.  Cost per block is independent of the other blocks.
# ------------------------------------------------------------------------------------------------
def cpu_sample(a, img, blocks_per_level=0, offset=0):
    """One bounded sample of the reference's CPU path: `m` range blocks per quadtree level (every (n/m)-th block of the level,
    shifted by `offset` so that successive samples take different blocks) against the FULL domain grid."""
    from oracle import pyoracle as po
    # the reference as its own CMake builds it: -mavx2 + FMA, -DFRAC_WITH_AVX=1 (oracle/Makefile: -march=x86-64-v3, portable to
    # the GPU box's host); all host threads (OpenMP over range blocks, TransformEstimator2::estimate per block)
    lib, kind = po.reference(fma=True), "reference"
    if lib is None:
        lib, kind = po.restatement(), "port"
    cores = lib.hardware_threads()
    m = blocks_per_level or 4 * cores
    W = H = a.size
    p = lib.params(a.thr, -1.0, bool(a.classifier), True)
    top = lib.uniform_grid(W, H, a.tmax, a.tmax)
    stride = max(1, len(top) // m)
    sample = top[offset % stride:: stride][:m]
    pending_est = float(len(top))
    tot_time = tot_matches = 0.0
    per_level = []
    T = a.tmax
    cpu_work = 0.0
    while T >= a.tmin and len(sample):
        dom = lib.uniform_grid(W, H, 2 * T, T)
        if a.classifier:
            dom, sample = lib.preclassify(img, dom), lib.preclassify(img, sample)
        t0 = time.perf_counter()
        out = lib.encode_level(img, img, dom, sample, p, 0, 1)
        dt = time.perf_counter() - t0
        cpu_work += dt
        if a.classifier:
            cand = float(np.mean([np.count_nonzero(dom["bin"] == b) for b in sample["bin"]])) * 4
        else:
            cand = len(dom) * 4.0
        tot_time += pending_est * dt / len(sample)
        tot_matches += pending_est * cand
        split = (out["distance"] > a.thr) if T // 2 >= a.tmin else np.zeros(len(out), bool)
        per_level.append({"T": T, "sampled": int(len(sample)), "sec": round(dt, 3), "split_frac": float(split.mean()),
                          "pending_est": pending_est})
        pending_est = pending_est * float(split.mean()) * 4
        kids = []
        for r in sample[split]:
            h = T // 2
            for dx, dy in ((0, 0), (h, 0), (0, h), (h, h)):
                kids.append((r["x"] + dx, r["y"] + dy, h, h, -1))
        sample = np.array(kids, po.GRID_ITEM) if kids else np.zeros(0, po.GRID_ITEM)
        if len(sample) > m:
            sample = sample[:: len(sample) // m][:m]
        T //= 2
    value = tot_matches / tot_time if tot_time > 0 else 0.0
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d range blocks per quadtree level vs the full domain grid, children of split sampled blocks feed the next level; "
                      "scaled per level by (estimated pending blocks)/(sampled blocks); %.1f s of CPU wall on %d threads (nproc %d); "
                      "build: g++ -O2 -march=x86-64-v3 -DFRAC_WITH_AVX=1 (AVX2+FMA, the reference's CMake flags), OpenMP over range blocks "
                      "around TransformEstimator2::estimate (EncodingEngineCore2's own thread pool deadlocks, SURVEY S9)" % (
                          m, cpu_work, cores, os.cpu_count() or cores),
            "cpu_work_s": cpu_work, "tot_matches": tot_matches,
            "est_image_seconds": tot_time, "est_mpix_per_s": W * H / 1e6 / tot_time if tot_time > 0 else 0.0, "levels": per_level}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    po.build(ref=os.path.isdir("/root/reference/encode"))
    fo = po.restatement()
    img = fo.synth_image(a.size, a.size, 1234, 0)
    res = []
    for step in range(a.warmup + a.steps):
        # every step samples DIFFERENT range blocks: the K timed steps together are one sample of K x m blocks per level
        r = cpu_sample(a, img, a.cpu_blocks, offset=step)
        if step >= a.warmup:
            res.append(r)
    # aggregate: whole-image time = mean of the per-step estimates; matches/s = nominal candidates / that time
    est_s = float(np.mean([r["est_image_seconds"] for r in res]))
    matches = float(np.mean([r["tot_matches"] for r in res]))
    v = matches / est_s if est_s > 0 else 0.0
    last = res[-1]
    work = float(sum(r["cpu_work_s"] for r in res))
    per_level_work = {}
    for r in res:
        for l in r["levels"]:
            per_level_work[l["T"]] = per_level_work.get(l["T"], 0.0) + l["sec"]
    sample = ("%d timed steps x (%s); each step takes different blocks: %.1f s of CPU wall in total, per level %s" % (
        len(res), last["sample"], work, ", ".join("T=%d %.1f s" % (t, w) for t, w in sorted(per_level_work.items(), reverse=True))))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": est_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32 (CPU AVX2)",
            "data": "synthetic", "config": {"workload": workload_name(a), "note": "whole-image time extrapolated from a bounded sample (cost per range "
                                            "block is independent of the other blocks); ms_per_step is that estimate; value = nominal candidates / it"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mpix_per_s": a.size * a.size / 1e6 / est_s if est_s > 0 else 0.0, "levels": last["levels"], "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        # an NVML query takes driver locks of the GPU it asks about: a few samples per timed region, not a busy poll
        self.interval = float(os.environ.get("FE_BENCH_CLOCK_INTERVAL", "0.1"))
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv or self.interval <= 0:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.interval)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def batch_config5(a, ctx, fb, rank, world, sync_all, size=1024, per_rank=32):
    """BASELINE config 5 in shape: a batch of 1024 x 1024 images, 8x8 grid + quadtree split to 4x4, image-parallel: every
    rank encodes `per_rank` images from pinned host memory through fe_encode_batch (two pipelined streams) into pinned
    host records.  Also the same images one call at a time (fe_encode_quadtree), to show what the pipelining buys."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    params = fb.Params(a.thr, -1.0, False, False, a.search)
    cap = (size // 4) ** 2
    host = torch.empty((per_rank, size, size), dtype=torch.uint8).pin_memory()
    for i in range(per_rank):
        ctx.set_synthetic_image(size, size, 5000 + rank * per_rank + i, 0)
        host[i].copy_(torch.from_numpy(ctx.get_image()))
    out = torch.empty(per_rank * cap * 64, dtype=torch.uint8).pin_memory()
    ptrs = (C.c_void_p * per_rank)(*[host[i].data_ptr() for i in range(per_rank)])
    counts = (C.c_size_t * per_rank)()

    def batched():
        rc = ctx.lib.fe_encode_batch(ctx.h, ptrs, per_rank, size, size, size, 8, 4, C.byref(params), out.data_ptr(), cap, counts)
        if rc != 0:
            raise fb.FractencodeError(rc, ctx.lib.fe_last_error(ctx.h).decode())

    def one_by_one():
        n = C.c_size_t(0)
        for i in range(per_rank):
            ctx.set_image(host[i].numpy())
            rc = ctx.lib.fe_encode_quadtree(ctx.h, 8, 4, C.byref(params), out.data_ptr() + i * cap * 64, cap, C.byref(n), None)
            if rc != 0:
                raise fb.FractencodeError(rc, ctx.lib.fe_last_error(ctx.h).decode())

    res = {}
    for name, fn in (("batched", batched), ("one_by_one", one_by_one)):
        fn()                                       # warm-up (allocations, slice hints)
        sync_all()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        v = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        res[name] = v.item()
    n_img = per_rank * world
    return {"workload": "%d natural %dx%d images per GPU (seeds 5000+), 8x8 grid + quadtree split to 4x4, full search, rms_threshold %g "
                        "(BASELINE config 5 in shape), pinned host images in, pinned host records out" % (per_rank, size, size, a.thr),
            "n_gpus": world, "images": n_img, "images_per_s": n_img / res["batched"], "ms_per_image": res["batched"] / per_rank * 1e3,
            "mpix_per_s": n_img * size * size / 1e6 / res["batched"],
            "one_call_per_image": {"images_per_s": n_img / res["one_by_one"], "ms_per_image": res["one_by_one"] / per_rank * 1e3},
            "timing": "host wall clock around the call(s) + device synchronise, max over ranks (the call is synchronous)"}


def strong_config4(a, ctx, fb, rank, world, sync_all, size=8192, steps=3):
    """BASELINE config 4 in the same run: ONE 8192 x 8192 image with Classifier2, its top-level range blocks sharded over the N
    GPUs (domain pool rebuilt on every GPU, per-rank lists gathered to rank 0), against the same image encoded by rank 0 alone."""
    import torch
    import torch.distributed as dist
    from fractencode_b200.dist import PackedGather, shard_slice
    params = fb.Params(a.thr, -1.0, True, False, a.search)
    ctx.set_synthetic_image(size, size, 4321, 0)
    n_top = (size // a.tmax) ** 2
    mine = shard_slice(n_top, rank, world)
    stream = torch.cuda.current_stream()
    per_top = (a.tmax // a.tmin) ** 2
    pg = PackedGather(-(-n_top // world) * per_top, torch.device("cuda"))      # the largest shard's worst case, equal on every rank

    def run(fn, n):
        ts = []
        for _ in range(n):
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return ts

    def sharded():
        n = ctx.encode_quadtree_slice_device(a.tmax, a.tmin, params, mine.start, mine.stop - mine.start)
        return pg.gather(ctx, n, a.tmax, shared_image=True)

    def whole():
        if rank == 0:
            ctx.encode_quadtree_device(a.tmax, a.tmin, params)

    run(sharded, 1)
    t_n = run(sharded, steps)
    run(whole, 1)
    t_1 = run(whole, steps)
    v = torch.tensor([float(np.mean(t_n)), float(np.mean(t_1))], dtype=torch.float64, device="cuda")
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    tn, t1 = v[0].item(), v[1].item()
    return {"workload": "natural %dx%d u8 seed 4321, quadtree %d->%d, Classifier2, rms_threshold %g (BASELINE config 4)" % (size, size, a.tmax, a.tmin, a.thr),
            "scaling": "strong", "n_gpus": world, "sharding": "top-level range blocks over the GPUs (fe_encode_quadtree_slice_device), packed 8-byte records "
            "gathered to rank 0", "ms_per_image_n_gpus": tn, "ms_per_image_1_gpu": t1, "speedup": t1 / tn if tn > 0 else None,
            "mpix_per_s": size * size / 1e6 / (tn * 1e-3), "steps": steps}


def run_b200(a):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import fractencode_b200 as fb
    from fractencode_b200.dist import PackedGather, shard_slice

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not os.path.exists(fb.library_path()):
        raise SystemExit("libfractencode_b200.so missing: run __graft_entry__.build() (no fallback path exists)")
    torch.cuda.set_device(local)
    if world > 1 and os.environ.get("FE_BENCH_PIN", "1") != "0" and hasattr(os, "sched_setaffinity"):
        # one slice of the host cores per rank: the level driver is a latency-bound host thread (one synchronisation per quadtree
        # level), and eight of them plus their NCCL / monitor threads migrating over each other show up as idle GPU time
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except OSError:
            pass
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    ctx = fb.Context(local, stream.cuda_stream)
    lib = ctx.lib
    W = H = a.size
    params = fb.Params(a.thr, -1.0, bool(a.classifier), False, a.search)
    cap = (W // a.tmin) * (H // a.tmin)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    host_img = torch.empty((H, W), dtype=torch.uint8).pin_memory()
    host_items = torch.empty(cap * 64, dtype=torch.uint8).pin_memory()

    by_ranges = a.shard == "ranges" and world > 1
    ctx.set_synthetic_image(W, H, 1234 + (0 if by_ranges else rank), 0)
    host_img.copy_(torch.from_numpy(ctx.get_image()))
    n_top = (W // a.tmax) * (H // a.tmax)
    mine = shard_slice(n_top, rank, world)          # this rank's top-level range blocks when sharding by ranges
    per_top = (a.tmax // a.tmin) ** 2
    pg = PackedGather(-(-n_top // world) * per_top if by_ranges else cap, torch.device("cuda")) if world > 1 else None

    def step_resident():
        if by_ranges:
            n = ctx.encode_quadtree_slice_device(a.tmax, a.tmin, params, mine.start, mine.stop - mine.start)
        else:
            n = ctx.encode_quadtree_device(a.tmax, a.tmin, params)
        if world > 1:  # gather the per-rank transform lists on rank 0 (the only collective of the path): packed 8-byte records
            pg.gather(ctx, n, a.tmax, shared_image=by_ranges)
        return n

    def step_e2e():
        ctx.set_image(host_img.numpy())  # pinned host -> device on the ctx stream
        n = C.c_size_t(0)
        if by_ranges:
            n = C.c_size_t(ctx.encode_quadtree_slice_device(a.tmax, a.tmin, params, mine.start, mine.stop - mine.start))
            rc = lib.fe_fetch_items(ctx.h, host_items.data_ptr(), cap, C.byref(n))
        else:
            rc = lib.fe_encode_quadtree(ctx.h, a.tmax, a.tmin, C.byref(params), host_items.data_ptr(), cap, C.byref(n), None)
        if rc != 0:
            raise fb.FractencodeError(rc, lib.fe_last_error(ctx.h).decode())
        return n.value

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        times = []
        n = 0
        for _ in range(steps):
            flush.fill_(1)  # evict L2 between timed iterations (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            n = fn()
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        return times, n

    for _ in range(a.warmup):
        step_resident()
    sync_all()
    ctx.stats_reset()
    sampler = ClockSampler(local)
    sampler.start()
    t_wall0 = time.perf_counter()
    times, n_items = timed(step_resident, a.steps)
    sync_all()
    wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    st = ctx.stats()
    launches = int(st.kernel_launches)
    matches_step = sum(int(x) for x in st.level_matches)
    evaluated_step = sum(int(x) for x in st.level_evaluated)
    nlev = int(np.log2(a.tmax // a.tmin)) + 1
    levels = []
    flops = 0.0
    search_ms = 0.0
    for l in range(nlev):
        T = a.tmax >> l
        lm, le, ms = int(st.level_matches[l]), int(st.level_evaluated[l]), float(st.level_search_ms[l])
        lp_ = int(st.level_prefiltered[l])
        lf = 2.0 * T * T * le + 2.0 * 64 * lp_      # prefilter pairs: an 8 x 8 bound, 64 products each
        flops += lf
        search_ms += ms
        levels.append({"T": T, "ranges": int(st.level_ranges[l]), "items": int(st.level_items[l]), "matches": lm, "evaluated": le, "prefiltered": lp_,
                       "passes": int(st.level_passes[l]), "search_ms": round(ms, 3), "prep_ms": round(float(st.level_prep_ms[l]), 3),
                       "tflops": round(lf / (ms * 1e-3) / 1e12, 2) if ms > 0 else None})
    # e2e through the C ABI with host buffers
    for _ in range(max(1, a.warmup // 2)):
        step_e2e()
    sync_all()
    e_times, e_items = timed(step_e2e, a.steps)
    sync_all()

    # The dominant search kernel at steady state: the same level scanned in ONE pass (no slices, no bins: every range block
    # against every domain, long work items) -- what the kernel does when the pruning of the step cannot shorten the scan.
    steady = None
    if rank == 0 and os.environ.get("FE_BENCH_STEADY", "1") != "0":
        dom_T = max(range(nlev), key=lambda l: float(st.level_search_ms[l]))
        Ts = a.tmax >> dom_T
        os.environ["FE_SINGLE_PASS"] = "1"
        try:
            for _ in range(2):
                ctx.stats_reset()
                ctx.encode_quadtree_device(Ts, Ts, params)
            s2 = ctx.stats()
            ms2, ev2 = float(s2.level_search_ms[0]), int(s2.level_evaluated[0])
            steady = {"T": Ts, "candidates": ev2, "search_ms": ms2, "achieved": 2.0 * Ts * Ts * ev2 / (ms2 * 1e-3) / 1e12 if ms2 > 0 else 0.0,
                      "note": "one-pass scan of the whole level (FE_SINGLE_PASS=1): all %d-pixel range blocks of the image x every domain x 4" % Ts}
        finally:
            del os.environ["FE_SINGLE_PASS"]

    # HBM-bound side of the path, measured once on rank 0: the decode gather (Decoder2's fixed-point iteration)
    decode_info = None
    if rank == 0:
        items_host = np.frombuffer(host_items.numpy()[: int(e_items) * 64].tobytes(), dtype=fb.ENCODE_ITEM)
        dec_iters = 8
        ctx.decode(items_host, W, H, max_iters=dec_iters, eps=-1e9)  # warm-up (allocations)
        ctx.decode(items_host, W, H, max_iters=dec_iters, eps=-1e9)
        dms = float(ctx.stats().last_decode_ms) / dec_iters
        dbytes = 2.0 * W * H + 64.0 * len(items_host)       # read plane + write plane + item records per iteration
        decode_info = {"ms_per_iteration": dms, "algorithmic_bytes": dbytes, "achieved_gbs": dbytes / (dms * 1e-3) / 1e9,
                       "note": "k_decode_step_small<8>/<4> (gather from the half-resolution box-sum plane the previous iteration wrote, convergence sum fused in: + W*H re-read of the old plane, + W*H/2 box sums written and read) + k_decode_check per iteration; the planes and the items stay in L2 at this size, the kernels are issue-bound (fp64 per pixel, as the reference computes) after the L1 wavefronts were cut (profiles/search_kernels_r2.md)"}

        if world == 1 and not a.no_decode_large and W * H < 8192 * 8192:
            # the same at 8192 x 8192, where the two planes (2 x 67 MB) and the items leave the 126 MB L2
            ctx.set_synthetic_image(8192, 8192, 4321, 0)
            big, _ = ctx.encode_quadtree(a.tmax, a.tmin, params)
            ctx.decode(big, 8192, 8192, max_iters=dec_iters, eps=-1e9)
            ctx.decode(big, 8192, 8192, max_iters=dec_iters, eps=-1e9)
            bms = float(ctx.stats().last_decode_ms) / dec_iters
            bbytes = 2.0 * 8192 * 8192 + 64.0 * len(big)
            decode_info["at_8192"] = {"ms_per_iteration": bms, "items": int(len(big)), "algorithmic_bytes": bbytes,
                                      "achieved_gbs": bbytes / (bms * 1e-3) / 1e9, "frac_hbm": bbytes / (bms * 1e-3) / 1e9 / peaks()["hbm_gbs"]}
            del big
            ctx.set_synthetic_image(W, H, 1234, 0)

    my = torch.tensor([sum(times), sum(e_times), float(matches_step), float(n_items)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = my.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = my.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tot_ms, e_ms, all_matches, all_items = tmax[0].item(), tmax[1].item(), tsum[2].item(), tsum[3].item()
        every = [torch.zeros_like(my) for _ in range(world)]
        dist.all_gather(every, my)
        per_rank_ms = [round(t[0].item() / a.steps, 3) for t in every]
    else:
        tot_ms, e_ms, all_matches, all_items = my[0].item(), my[1].item(), my[2].item(), my[3].item()
        per_rank_ms = [round(tot_ms / a.steps, 3)]
    strong = strong_config4(a, ctx, fb, rank, world, sync_all) if (world > 1 and not by_ranges and not a.no_strong) else None
    batch = batch_config5(a, ctx, fb, rank, world, sync_all) if (not by_ranges and not a.no_batch) else None
    if rank == 0:
        pk = peaks()
        ms_per_step = tot_ms / a.steps
        value = all_matches / (ms_per_step * 1e-3)
        e_value = all_matches / (e_ms / a.steps * 1e-3)
        ach = flops / (search_ms * 1e-3) / 1e12 if search_ms > 0 else 0.0
        # dominant kernel = the search kernel of the level with the largest share of the step
        dom = max(levels, key=lambda l: l["search_ms"])
        dom_ach = 2.0 * dom["T"] ** 2 * dom["evaluated"] / (dom["search_ms"] * 1e-3) / 1e12 if dom["search_ms"] > 0 else 0.0
        # Peak: MEASURED_PEAKS.json holds a burst figure (best of 10 single GEMMs) and a sustained one (back to back for 4 s).
        # A timed region of a fraction of a second at maximum clock is burst conditions; both fractions are printed.
        burst = wall < 2.0
        peak = pk["tflops_burst"] if burst else pk["tflops_sustained"]
        for l in levels:
            if l["tflops"]:
                l["frac_burst"] = round(l["tflops"] / pk["tflops_burst"], 3)
        traffic, traffic_src = measured_traffic(a, dom["T"])
        kname = "k_search_f16<%d>" % dom["T"] if dom["T"] <= 8 else "k_search_i8"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if by_ranges else "weak", "vs_baseline": None,
            "dtype": "u8 in; T<=8: fp16 operands, fp32 (integer-exact) accumulate; T>=16: u8 x u8 -> s32; int32 scores, f64 s/o",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "images_per_step": 1 if by_ranges else world,
                       "sharding": "top-level range blocks of one image over the GPUs, pool rebuilt per GPU" if by_ranges else "one image per GPU",
                       "l2": "flushed between timed steps (512 MiB fill)",
                       "search_impl": ["auto", "exact-int (dp4a)", "tcgen05"][a.search], "pruning": PRUNING_NOTE,
                       "value_counts": "NOMINAL candidates (SURVEY 8d: the full candidate count of the workload, what the reference scores without "
                                       "its break) per second -- a time ratio, not device work; evaluated_matches_per_s and mpix_per_s are the "
                                       "device-work and end-user figures"},
            "mpix_per_s": (1 if by_ranges else world) * W * H / 1e6 / (ms_per_step * 1e-3),
            "evaluated_matches_per_s": evaluated_step * (1 if by_ranges else world) / (ms_per_step * 1e-3),
            "e2e": {"value": e_value, "unit": UNIT, "h2d_bytes_per_step": W * H, "d2h_bytes_per_step": int(e_items) * 64,
                    "ms_per_step": e_ms / a.steps, "mpix_per_s": (1 if by_ranges else world) * W * H / 1e6 / (e_ms / a.steps * 1e-3)},
            "gpu_launches": launches,
            "per_rank_ms_per_step": per_rank_ms, "host_cores": os.cpu_count(),
            "items_per_step": all_items,
            "evaluated_matches_per_step": evaluated_step,
            "levels": levels,
            "umma_levels": int(st.umma_levels), "exact_levels": int(st.exact_levels),
            "roofline": {"bound": "tensor", "achieved": dom_ach, "peak": peak, "unit": "TFLOP/s", "frac": dom_ach / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_kind": "burst" if burst else "sustained",
                         "frac_burst": dom_ach / pk["tflops_burst"], "frac_sustained": dom_ach / pk["tflops_sustained"],
                         "kernel": "%s, level T=%d: 2*T^2 FLOP x %d candidates scored by the level's %d search launches / their summed CUDA-event "
                                   "durations (ctx stream); peak = %s bf16 dense (timed region %.2f s), %s" % (
                             kname, dom["T"], dom["evaluated"], dom["passes"], "burst" if burst else "sustained", wall, pk["source"]),
                         "all_levels": {"achieved": ach, "frac_burst": ach / pk["tflops_burst"], "frac_sustained": ach / pk["tflops_sustained"]},
                         "steady_state": dict(steady, frac_burst=steady["achieved"] / pk["tflops_burst"],
                                              frac_sustained=steady["achieved"] / pk["tflops_sustained"]) if steady else None,
                         "peak_burst": pk["tflops_burst"], "peak_sustained": pk["tflops_sustained"]},
            "hbm_kernels": {"peak_gbs": pk["hbm_gbs"], "decode": decode_info,
                            "prep_ms_per_level": {str(l["T"]): l["prep_ms"] for l in levels}},
            "clocks": sampler.summary(),
            "wall_s_timed_region": wall,
        }
        if strong:
            line["strong"] = strong
        if batch:
            line["batch"] = batch
        if world == 1 and not a.no_cpu_baseline:
            from oracle import pyoracle as po
            po.build(ref=os.path.isdir("/root/reference/encode"))
            cb = cpu_sample(a, host_img.numpy(), a.cpu_blocks)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["est_mpix_per_s"] = cb["est_mpix_per_s"]
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
