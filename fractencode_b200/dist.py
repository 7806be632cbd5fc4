"""Multi-GPU plumbing of the path: shard independent units (images / range blocks) over ranks, gather the
per-rank transform lists.  The search itself needs no collective: the domain pool is rebuilt locally
from the replicated image.  torch.distributed supplies the process group (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

ITEM_BYTES = 64


def shard_slice(n_units: int, rank: int, world: int) -> slice:
    """Contiguous, balanced slice of n_units for this rank (first n_units % world ranks get one more)."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def gather_item_lists(items_u8: torch.Tensor, n_items: int, cap_items: int, counts_out: torch.Tensor | None = None,
                      gather_out: torch.Tensor | None = None, group=None):
    """All-gather variable-length lists of 64-byte encode_item_t records.

    items_u8: this rank's records as a flat uint8 tensor of at least cap_items*64 bytes (device or CPU);
    n_items:  valid records on this rank; cap_items: common capacity of the buffers.
    Returns (counts [world] int64, gathered [world, row_items*64] uint8) with row_items = the largest count (rounded up
    to 1024 records): one small all-gather for the counts -- read back on the host to size the second -- and one
    all-gather of equal-sized record blocks.  Gathering the full capacity instead would move 8 x 67 MB per 4096^2 image
    on 8 GPUs for 21 MB of records each."""
    world = dist.get_world_size(group)
    dev = items_u8.device
    if counts_out is None:
        counts_out = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n_items], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_out, mine, group=group)
    row_items = min(cap_items, (int(counts_out.max().item()) + 1023) // 1024 * 1024)
    row_items = max(row_items, 1)
    if gather_out is None:
        gather_out = torch.empty(world * row_items * ITEM_BYTES, dtype=torch.uint8, device=dev)
    out = gather_out[: world * row_items * ITEM_BYTES]
    dist.all_gather_into_tensor(out, items_u8[: row_items * ITEM_BYTES].contiguous(), group=group)
    return counts_out, out.view(world, row_items * ITEM_BYTES)


def unpack_gathered(counts: torch.Tensor, gathered: torch.Tensor, dtype: np.dtype) -> list[np.ndarray]:
    """Host view: one structured array per rank, trimmed to its count."""
    out = []
    g = gathered.cpu().numpy()
    for r, n in enumerate(counts.cpu().tolist()):
        out.append(np.frombuffer(g[r].tobytes(), dtype=dtype, count=n).copy())
    return out
