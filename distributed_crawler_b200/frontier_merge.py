"""Multi-GPU frontier set merge (option A: hash-partitioned owners).

The record path shards on record index with no data-path collective; only the dedup set is global.  After the
local pass every rank holds an exact local set of 32-byte keys in first-occurrence order.  One merge = ONE exchange
step: the keys a rank added since the last merge are bucketed by owner = hash(key) % world, every bucket goes to its
owner, and the owner inserts what it received (source-rank-major, so the lowest rank's copy of a key wins) into its
partition of the global set.  A key carries the sequence number (merge round, source rank, position in the source's
set) of its first occurrence: the union of the partitions ordered by that number is exactly the set one process
would have built over the ranks' shards in rank order.  Per-rank insert work is O(new keys / world).

Two implementations of the same protocol:
  * `NcclMerger` — the product path: `tgi_comm_init` / `tgi_frontier_merge` in libtgingest (device-side bucketing,
    grouped ncclSend / ncclRecv over NVLink, owner-side insert kernels); torch.distributed only carries the 128-byte
    NCCL id from rank 0 to the other ranks, as a Go host would do over its own control channel.
  * `PartitionedMerge` — the protocol restated over torch.distributed objects with any `KeySet` (CPU): the host-logic
    double that the world_size-2 gloo tests run; it is never used on a GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def key_owner(keys: np.ndarray, world: int) -> np.ndarray:
    """owner rank of every 32-byte key (any deterministic hash works for the CPU double; the device code uses its own)"""
    k = np.ascontiguousarray(keys, np.uint8).reshape(-1, 32).view("<u8")
    h = np.full(len(k), 0x9E3779B97F4A7C15, np.uint64)
    with np.errstate(over="ignore"):
        for j in range(4):
            h = (h ^ k[:, j]) * np.uint64(0xFF51AFD7ED558CCD)
            h ^= h >> np.uint64(32)
    return ((h >> np.uint64(17)) % np.uint64(world)).astype(np.int64)


class PartitionedMerge:
    """The merge protocol over torch.distributed collectives of Python objects (gloo on CPU).

    `local` and `owned` are KeySets: `export_new(first) -> np.ndarray[m,32]`, `size()`, `insert(keys) -> is_new mask`.
    """

    def __init__(self, local, owned, group=None):
        self.local, self.owned, self.group = local, owned, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.merged_upto = 0
        self.round = 0
        self.payload = []  # sequence number of every key of the owned partition, in insertion order
        self.sent = 0

    def merge(self) -> int:
        first = self.merged_upto
        mine = np.ascontiguousarray(self.local.export_new(first), np.uint8).reshape(-1, 32)
        own = key_owner(mine, self.world) if len(mine) else np.zeros(0, np.int64)
        seq = (self.round << 52) | (self.rank << 44) | (first + np.arange(len(mine), dtype=np.int64))
        buckets = [(mine[own == p], seq[own == p]) for p in range(self.world)]
        self.sent += sum(len(b[0]) for p, b in enumerate(buckets) if p != self.rank)
        # the exchange: rank q receives bucket q of every rank, laid out by source rank
        recv = [None] * self.world
        for p in range(self.world):
            got = [None] * self.world if self.rank == p else None
            dist.gather_object(buckets[p], got, dst=p, group=self.group)
            if self.rank == p:
                recv = got
        for keys, seqs in recv:
            if len(keys):
                new = np.asarray(self.owned.insert(keys), bool)
                self.payload += [int(s) for s in seqs[new]]
        sizes = [None] * self.world
        dist.all_gather_object(sizes, self.owned.size(), group=self.group)
        self.merged_upto = first + len(mine)
        self.round += 1
        return int(sum(sizes))

    def global_export(self) -> np.ndarray:
        """every partition, ordered by first-occurrence sequence number, on every rank"""
        part = (np.ascontiguousarray(self.owned.export_new(0), np.uint8).reshape(-1, 32), np.asarray(self.payload, np.int64))
        parts = [None] * self.world
        dist.all_gather_object(parts, part, group=self.group)
        keys = np.concatenate([p[0] for p in parts]) if parts else np.zeros((0, 32), np.uint8)
        seqs = np.concatenate([p[1] for p in parts])
        return keys[np.argsort(seqs, kind="stable")]


class NcclMerger:
    """libtgingest's own merge (tgi_comm_init / tgi_frontier_merge) for one Engine per rank."""

    def __init__(self, engine, device: torch.device, group=None):
        self.e = engine
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        uid = torch.zeros(128, dtype=torch.uint8, device=device if dist.get_backend(group) == "nccl" else "cpu")
        if self.rank == 0:
            uid.copy_(torch.frombuffer(bytearray(engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0, group=group)
        engine.comm_init(bytes(uid.cpu().numpy().tobytes()), self.rank, self.world)

    def merge(self) -> int:
        return self.e.frontier_merge()[0]

    def global_export(self) -> np.ndarray:
        return self.e.frontier_global_export()

    def stats(self) -> dict:
        return self.e.merge_stats()

    @staticmethod
    def describe() -> str:
        return ("tgi_frontier_merge: new keys bucketed on the device by owner = hash % ranks, one ncclAllGather of the counts, "
                "grouped ncclSend/ncclRecv of the buckets, owner-side insert, ncclAllReduce of the partition sizes")


def make_merger(engine, device: torch.device, group=None) -> NcclMerger:
    return NcclMerger(engine, device, group)
