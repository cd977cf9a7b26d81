"""The CPU arm's parallel driver (oracle/tgoracle.c run_batch: produce / place / sharded frontier insert) must return
exactly what the sequential reference loop returns — the timed baseline is only meaningful if it is the same work."""
import numpy as np
import pytest

from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus, YtCorpus
from oracle.pyoracle import Oracle

ALL = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
ORC_RUN_SLICES, ORC_RUN_PIN = 0x10000, 0x20000


def _same(a, b):
    for k in ("status", "jsonl", "line_off", "link_off", "links"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert (a.n_new, a.frontier_size) == (b.n_new, b.frontier_size)


@pytest.mark.parametrize("nthreads", [2, 5, 8])
def test_parallel_equals_sequential_across_batches(nthreads):
    o1, on = Oracle(), Oracle()
    for rnd, (n, prof) in enumerate([(30_000, 2), (20_000, 3), (7, 3)]):
        c = Corpus(n, seed=0x5EED0002 + rnd, first=rnd * 1000, profile=prof, nthreads=2)
        _same(o1.telegram(c.batch, ALL, nthreads=1), on.telegram(c.batch, ALL, nthreads=nthreads))
        assert np.array_equal(o1.frontier_export(), on.frontier_export())  # content AND first-occurrence order


def test_tandem_filter_and_youtube_parallel():
    o1, on = Oracle(), Oracle()
    c = Corpus(15_000, profile=3, nthreads=2)
    f = abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_FILTER | abi.RUN_SKIP_SELF
    _same(o1.telegram(c.batch, f, nthreads=1), on.telegram(c.batch, f, nthreads=7))
    y = YtCorpus(5_000, nthreads=2)
    fy = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER
    _same(o1.youtube(y.batch, fy, nthreads=1), on.youtube(y.batch, fy, nthreads=6))
    assert np.array_equal(o1.frontier_export(), on.frontier_export())


def test_slices_mode_keeps_everything_but_the_concatenation():
    c = Corpus(10_000, profile=2, nthreads=2)
    a, b = Oracle(), Oracle()
    ra = a.telegram(c.batch, ALL, nthreads=4)
    n, jl, nl = b.telegram(c.batch, ALL | ORC_RUN_SLICES | ORC_RUN_PIN, nthreads=4, copy=False)
    assert (n, jl, nl) == (ra.n, len(ra.jsonl), len(ra.links))
    assert np.array_equal(a.frontier_export(), b.frontier_export())
