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
    from fractencode_b200.dist import HEADER_WORDS, PackedGather, shard_slice, split_gathered

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 37 range blocks sharded over the ranks, every rank "encodes" its slice into fake packed records
    units = np.arange(37)
    sl = shard_slice(len(units), rank, world)
    mine = (units[sl].astype(np.uint64) << np.uint64(22)) | np.uint64(rank)
    pg = PackedGather(64, torch.device("cpu"))
    pg.send[HEADER_WORDS: HEADER_WORDS + len(mine)] = torch.from_numpy(mine.view(np.int64).copy())
    # per-rank min/max of (s, o), reduced like the Quantizer header of one image sharded over the ranks
    pg.send[1:5] = torch.tensor([-1.0 - rank, 2.0 + rank, 10.0 * (rank + 1), 50.0 - rank], dtype=torch.float64).view(torch.int64)
    pg.reduce_minmax()
    got = pg.exchange(len(mine))
    if rank == 0:
        lists = split_gathered(got)
        assert [len(p) for p, _ in lists] == [19, 18]
        allrec = np.concatenate([p for p, _ in lists])
        assert ((allrec >> np.uint64(22)) == units.astype(np.uint64)).all()
        assert [int(p[0] & np.uint64(1)) for p, _ in lists] == [0, 1]
        assert lists[0][1].tolist() == [-2.0, 3.0, 10.0, 50.0]        # min_s, max_s, min_o, max_o over both ranks
        assert tuple(got.shape) == (2, HEADER_WORDS + 64)
    else:
        assert got is None
    # slices tile the unit range exactly
    sls = [shard_slice(37, r, world) for r in range(world)]
    assert sls[0].start == 0 and sls[-1].stop == 37 and all(sls[i].stop == sls[i + 1].start for i in range(world - 1))
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
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "matches/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0
