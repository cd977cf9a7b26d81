"""world_size-2 gloo test (CPU) of the multi-GPU frontier merge protocol (SURVEY 8e option A): PartitionedMerge,
the host-logic double of tgi_frontier_merge, driven with CPU sets (the oracle's frontier, test infrastructure).
On GPUs the same protocol runs inside libtgingest over NCCL (tests/test_gpu_merge.py)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.frontier_merge import PartitionedMerge, key_owner

FLAGS = abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF


class OracleSet:
    def __init__(self):
        from oracle.pyoracle import Oracle
        self.o = Oracle()

    def size(self):
        return len(self.o.frontier_export())

    def export_new(self, first):
        return self.o.frontier_export()[first:].copy()

    def insert(self, keys):
        return self.o.frontier_insert(np.ascontiguousarray(keys))


def _worker(rank, world, port, n_per, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local, owned = OracleSet(), OracleSet()
    pm = PartitionedMerge(local, owned)
    sizes = []
    for rnd in range(2):  # two batches per rank with a merge after each
        c = Corpus(n_per, first=(rnd * world + rank) * n_per, profile=3, nthreads=1)  # record-index sharding
        local.o.telegram(c.batch, FLAGS)
        sizes.append(pm.merge())
    again = pm.merge()  # nothing new anywhere -> idempotent
    exp = pm.global_export()
    part = owned.export_new(0)
    q.put((rank, sizes, again, exp.tobytes(), bool(len(part) == 0 or (key_owner(part, world) == rank).all()), local.size(), pm.sent))
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_merge_two_ranks_equals_single_process():
    n_per, world = 4000, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: ONE process over the shards in (round, rank) order = the concatenated record range
    from oracle.pyoracle import Oracle
    o = Oracle()
    want_sizes = []
    for rnd in range(2):
        o.telegram(Corpus(n_per * world, first=rnd * world * n_per, profile=3, nthreads=2).batch, FLAGS)
        want_sizes.append(len(o.frontier_export()))
    want = o.frontier_export()
    for rank, sizes, again, exp, owner_ok, local_size, sent in out:
        assert sizes == want_sizes == [sizes[0], again], f"rank {rank}: global sizes {sizes} / {again}, want {want_sizes}"
        assert exp == want.tobytes(), f"rank {rank}: merged set (content or first-occurrence order) differs from the single-process set"
        assert owner_ok, "a key sits in a partition that does not own it"
        assert 0 < sent < local_size  # only the other rank's bucket travels
    assert out[0][5] + out[1][5] >= len(want)  # the shards overlap in names, the union dedups
