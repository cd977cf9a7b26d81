"""Multi-GPU frontier set merge (option B).

The record path shards on record index with no data-path collective; only the dedup set is global.
After the local pass every rank holds an exact local set of 32-byte keys.  The merge is ONE exchange
step: all-gather the keys each rank added since the last merge (NCCL over NVLink on GPUs) and insert
the gathered keys into the local set, so that every rank ends with the same global set.  First-
occurrence order is preserved within a rank; across ranks the merged order is rank-major.

The set itself is abstracted (`export_new(first)` / `insert(keys)`): on GPUs it is the device-resident
frontier of libtgingest (device pointers go straight into the collective, no host staging); the
world_size-2 gloo tests drive the same exchange logic with a CPU set.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist


class EngineFrontier:
    """Adapter over libtgingest's device-resident frontier (keys never leave HBM)."""

    def __init__(self, engine, device: torch.device):
        from . import engine as _e
        self._lib = _e.lib()
        self.e = engine
        self.device = device

    def size(self) -> int:
        return self.e.frontier_size()

    def export_new(self, first: int) -> torch.Tensor:
        n = max(self.size() - first, 0)
        out = torch.empty((n, 32), dtype=torch.uint8, device=self.device)
        if n:
            got = C.c_uint64()
            torch.cuda.synchronize(self.device)  # `out` was allocated on torch's stream
            self.e._check(self._lib.tgi_frontier_export_dev(self.e.h, out.data_ptr(), n, first, C.byref(got)))
            assert got.value == n
        return out

    def insert(self, keys: torch.Tensor) -> int:
        n = int(keys.shape[0])
        if n:
            keys = keys.contiguous()
            # the library launches on its own stream: the collective that produced `keys` (torch /
            # NCCL streams) must have finished before its kernels read them
            torch.cuda.synchronize(self.device)
            self.e._check(self._lib.tgi_frontier_insert_dev(self.e.h, keys.data_ptr(), n, None))
        return self.size()


def merge_frontier(fset, merged_upto: int, group=None) -> tuple[int, int]:
    """Exchange the keys added locally since `merged_upto` and insert everyone else's.
    Returns (global_size, new merged_upto).  Collective: every rank must call it."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = fset.export_new(merged_upto)
    dev = mine.device
    cnt = torch.tensor([mine.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts) if counts else 0
    if mx == 0:
        return fset.size(), fset.size()
    padded = torch.zeros((mx, 32), dtype=torch.uint8, device=dev)
    padded[: mine.shape[0]] = mine
    gathered = [torch.empty((mx, 32), dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    others = [gathered[r][: counts[r]] for r in range(world) if r != rank and counts[r]]
    if others:
        fset.insert(torch.cat(others, 0))
    size = fset.size()
    return size, size
