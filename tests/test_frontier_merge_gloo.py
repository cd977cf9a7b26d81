"""world_size-2 gloo test (CPU) of the multi-GPU frontier merge: the exchange logic of
distributed_crawler_b200/frontier_merge.py driven with a CPU set (the oracle's frontier, which is
test infrastructure).  On GPUs the same function runs over NCCL with libtgingest's device set."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.frontier_merge import merge_frontier


class OracleFrontier:
    def __init__(self):
        from oracle.pyoracle import Oracle
        self.o = Oracle()

    def size(self):
        return len(self.o.frontier_export())

    def export_new(self, first):
        return torch.from_numpy(self.o.frontier_export()[first:].copy())

    def insert(self, keys):
        self.o.frontier_insert(keys.numpy())
        return self.size()


def _worker(rank, world, port, n_per, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fs = OracleFrontier()
    c = Corpus(n_per, first=rank * n_per, profile=3, nthreads=1)  # record-index sharding
    r = fs.o.telegram(c.batch, abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF)
    local = fs.size()
    gsize, upto = merge_frontier(fs, 0)
    keys = sorted(bytes(k) for k in fs.o.frontier_export())
    # second round: nothing new anywhere -> idempotent
    g2, _ = merge_frontier(fs, upto)
    q.put((rank, local, gsize, g2, keys, int(r.n_new)))
    dist.barrier()
    dist.destroy_process_group()


def test_merge_two_ranks_equals_single_process():
    n_per, world = 6000, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: one process over the concatenated range
    from oracle.pyoracle import Oracle
    o = Oracle()
    o.telegram(Corpus(n_per * world, profile=3, nthreads=2).batch, abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF)
    want = sorted(bytes(k) for k in o.frontier_export())
    for rank, local, gsize, g2, keys, n_new in out:
        assert keys == want, f"rank {rank}: merged set differs from the single-process set"
        assert gsize == len(want) == g2
        assert local == n_new <= gsize
    assert out[0][1] + out[1][1] >= len(want)  # the shards overlap in names, the union dedups
