"""Multi-GPU plumbing of the path: shard independent units (images / range blocks) over ranks, gather the
per-rank transform lists.  The search itself needs no collective: the domain pool is rebuilt locally
from the replicated image.  torch.distributed supplies the process group (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_slice(n_units: int, rank: int, world: int) -> slice:
    """Contiguous, balanced slice of n_units for this rank (first n_units % world ranks get one more)."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


# ---------------------------------------------------------------------------------------------------------------
# Packed gather: the per-rank lists travel as 8-byte quantised records (include/fractencode_b200.h), only to the root,
# with the count and the Quantizer header in-band -- no host read-back anywhere on the way.
# ---------------------------------------------------------------------------------------------------------------
HEADER_WORDS = 8          # word 0: record count; words 1-4: min_s, max_s, min_o, max_o (f64 bit patterns); 5-7 reserved


class PackedGather:
    """Buffers and steps of one rank.  `cap_items` = the rank's worst-case record count (all blocks at t_min)."""

    def __init__(self, cap_items: int, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.cap = int(cap_items)
        self.words = HEADER_WORDS + self.cap
        self.send = torch.zeros(self.words, dtype=torch.int64, device=device)
        self.recv = torch.zeros((self.world, self.words), dtype=torch.int64, device=device) if self.rank == 0 else None
        self.red = torch.zeros(4, dtype=torch.float64, device=device)
        self.count_host = torch.zeros(1, dtype=torch.int64)
        if torch.device(device).type == "cuda":
            self.count_host = self.count_host.pin_memory()      # the count goes up with an asynchronous copy, no host stall

    def gather(self, ctx, n_items: int, t_max: int, shared_image: bool, bits=(5, 7)):
        """Pack this rank's device-resident result list and gather all lists on rank 0.

        shared_image: the ranks hold shards of ONE image -> the Quantizer's min/max run over the whole list (main.cpp:109-118):
        one 4-double all-reduce.  Otherwise (one image per rank) every list is quantised with its own header.
        Returns the [world, words] int64 tensor on rank 0 (None elsewhere); parse with `split_gathered`."""
        mm = self.send[1:5].view(torch.float64)
        ctx.items_minmax_device(mm.data_ptr())
        if shared_image:
            self.reduce_minmax()
        ctx.pack_items_device(t_max, mm.data_ptr(), self.send[HEADER_WORDS:].data_ptr(), self.cap, bits[0], bits[1])
        return self.exchange(n_items)

    def reduce_minmax(self):
        """min / max of the four header values over the ranks (in place, one all-reduce)."""
        if self.world == 1:
            return
        mm = self.send[1:5].view(torch.float64)
        self.red.copy_(torch.stack([mm[0], -mm[1], mm[2], -mm[3]]))
        dist.all_reduce(self.red, op=dist.ReduceOp.MIN, group=self.group)
        mm.copy_(torch.stack([self.red[0], -self.red[1], self.red[2], -self.red[3]]))

    def exchange(self, n_items: int):
        """Ship `send` (header + records) to rank 0; the count travels in word 0."""
        self.count_host[0] = int(n_items)
        self.send[0:1].copy_(self.count_host, non_blocking=True)
        if self.world == 1:
            return self.send.view(1, -1)
        dist.gather(self.send, [self.recv[r] for r in range(self.world)] if self.rank == 0 else None, dst=0, group=self.group)
        return self.recv


def split_gathered(gathered: torch.Tensor):
    """Host view of PackedGather.gather's result: [(packed uint64 array, minmax float64[4])] per rank."""
    g = gathered.cpu().numpy()
    out = []
    for row in g:
        n = int(row[0])
        out.append((row[HEADER_WORDS: HEADER_WORDS + n].view(np.uint64).copy(), row[1:5].view(np.float64).copy()))
    return out
