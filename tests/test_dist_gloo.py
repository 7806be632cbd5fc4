"""N > 1 host logic on CPU: world_size-2 gloo run of the shard + gather path bench.py uses with NCCL."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from fractencode_b200.capi import ENCODE_ITEM
    from fractencode_b200.dist import gather_item_lists, shard_slice, unpack_gathered

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 37 range blocks sharded over the ranks, every rank "encodes" its slice into fake records
    units = np.arange(37)
    sl = shard_slice(len(units), rank, world)
    mine = np.zeros(sl.stop - sl.start, ENCODE_ITEM)
    mine["x"] = units[sl] * 4
    mine["y"] = rank
    mine["distance"] = units[sl] * 0.5
    cap = 2048
    buf = torch.zeros(cap * 64, dtype=torch.uint8)
    buf[: mine.nbytes] = torch.from_numpy(np.frombuffer(mine.tobytes(), np.uint8).copy())
    counts, gathered = gather_item_lists(buf, len(mine), cap)
    lists = unpack_gathered(counts, gathered, ENCODE_ITEM)
    allitems = np.concatenate(lists)
    assert counts.tolist() == [19, 18], counts
    assert tuple(gathered.shape) == (2, 1024 * 64), gathered.shape   # blocks of the largest count (rounded), not of the capacity
    assert (allitems["x"] == units * 4).all()
    assert (allitems["distance"] == units * 0.5).all()
    assert [int(l["y"][0]) for l in lists] == [0, 1]
    # slices tile the unit range exactly
    got = [shard_slice(37, r, world) for r in range(world)]
    assert got[0].start == 0 and got[-1].stop == 37 and all(got[i].stop == got[i + 1].start for i in range(world - 1))
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def test_shard_and_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("ok") == 2


def test_bench_reference_arm_small():
    """bench.py --impl reference prints one well-formed JSON line (bounded sample of the workload, CPU only)."""
    import json
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "256", "--tmax", "16", "--steps", "1",
                        "--warmup", "0", "--cpu-blocks", "4"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().split("\\n")[-1])
    assert line["impl"] == "reference" and line["unit"] == "matches/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0
