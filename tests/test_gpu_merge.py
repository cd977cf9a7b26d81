"""Multi-GPU dedup-set merge ON the GPUs (needs >= 2 devices, else skipped): one process per GPU, libtgingest's own
NCCL merge (tgi_comm_init / tgi_frontier_merge); the merged global set must equal — content AND first-occurrence
order — the set the single-process oracle builds over the same record range."""
import os
import socket

import numpy as np
import pytest
import torch

from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus

pytestmark = pytest.mark.gpu
FLAGS = abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
ROUNDS = 2


def _worker(rank, world, port, n_per, q):
    import torch.distributed as dist
    from distributed_crawler_b200.engine import Engine
    from distributed_crawler_b200.frontier_merge import make_merger
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # carries the NCCL id only
    torch.cuda.set_device(rank)
    e = Engine(device=rank, frontier_capacity=1 << 22)
    m = make_merger(e, torch.device("cuda", rank))
    sizes = []
    for rnd in range(ROUNDS):
        c = Corpus(n_per, first=(rnd * world + rank) * n_per, profile=3, nthreads=4)
        e.telegram(c.batch, FLAGS, copy=False)
        sizes.append(m.merge())
    again = m.merge()
    exp = m.global_export()
    st = m.stats()
    q.put((rank, sizes, again, exp.tobytes(), st["keys_sent"], st["keys_owned"], e.frontier_size()))
    dist.barrier()
    e.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_merge_equals_single_process_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    n_per = 300_000
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from oracle.pyoracle import Oracle
    o = Oracle()
    want_sizes = []
    for rnd in range(ROUNDS):
        o.telegram(Corpus(n_per * world, first=rnd * world * n_per, profile=3).batch, FLAGS, nthreads=os.cpu_count() or 1, copy=False)
        want_sizes.append(len(o.frontier_export()))
    want = o.frontier_export().tobytes()
    owned_total = 0
    for rank, sizes, again, exp, sent, owned, local in out:
        assert sizes == want_sizes and again == want_sizes[-1], f"rank {rank}: {sizes} / {again} vs {want_sizes}"
        assert exp == want, f"rank {rank}: merged set (content or order) differs from the single-process oracle set"
        assert 0 < sent < local
        owned_total += owned
    assert owned_total == want_sizes[-1]  # the partitions are disjoint and cover the set


def test_single_rank_comm_is_a_plain_set():
    """world = 1: the merge degenerates to copying the local set into the (only) partition"""
    from distributed_crawler_b200.engine import Engine
    from oracle.pyoracle import Oracle
    e, o = Engine(), Oracle()
    e.comm_init(Engine.comm_unique_id(), 0, 1)
    c = Corpus(40_000, profile=3)
    e.telegram(c.batch, FLAGS, copy=False)
    o.telegram(c.batch, FLAGS)
    g, owned = e.frontier_merge()
    assert g == owned == len(o.frontier_export())
    assert np.array_equal(e.frontier_global_export(), o.frontier_export())
    e.close()
