"""GPU parity at the BASELINE.json sizes (configs[1..3] = bench.py --config 2/3/4): the CUDA path against the CPU
oracle on the SAME full-size corpora the bench uses — every status byte, every JSONL byte, every per-record link and the
frontier (content and first-occurrence order).  TGI_TEST_SCALE < 1 (or a host with little RAM) shrinks the corpora."""
import os

import numpy as np
import pytest

from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus, YtCorpus
from distributed_crawler_b200.engine import Engine
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 1


def _scale():
    s = float(os.environ.get("TGI_TEST_SCALE", "1"))
    try:
        ram = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
        if ram < 300e9:
            s = min(s, ram / 300e9)
    except (ValueError, OSError):
        pass
    if CORES < 32:
        s = min(s, CORES / 32)
    return max(s, 0.002)


def _view(p, n, dt):
    import ctypes as C
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (n * np.dtype(dt).itemsize,)).view(dt) if n else np.zeros(0, dt)


def _compare_batch(e, o, batch, flags, kind, label):
    """one batch through both paths without copying the big arrays more than once"""
    import ctypes as C
    from oracle import pyoracle
    d = batch.descriptor()
    ro = pyoracle.OrcResultC()
    fn_o = pyoracle.lib().orc_youtube_batch if kind == "yt" else pyoracle.lib().orc_telegram_batch
    assert fn_o(o.h, C.byref(d), flags, min(CORES, 128), C.byref(ro)) == 0
    rg = abi.ResultC()
    from distributed_crawler_b200 import engine as eng_mod
    fn_g = eng_mod.lib().tgi_youtube_batch if kind == "yt" else eng_mod.lib().tgi_telegram_batch
    e._check(fn_g(e.h, C.byref(d), flags, C.byref(rg)))
    try:
        n = int(ro.n)
        assert n == int(rg.n) == batch.n
        assert np.array_equal(_view(ro.status, n, np.uint8), _view(rg.status, n, np.uint8)), f"{label}: status"
        if flags & abi.RUN_JSONL:
            assert int(ro.jsonl_len) == int(rg.jsonl_len), f"{label}: JSONL length {ro.jsonl_len} vs {rg.jsonl_len}"
            assert np.array_equal(_view(ro.line_off, n + 1, np.uint64), _view(rg.line_off, n + 1, np.uint64)), f"{label}: line offsets"
            a, b = _view(ro.jsonl, int(ro.jsonl_len), np.uint8), _view(rg.jsonl, int(rg.jsonl_len), np.uint8)
            step = 1 << 30
            for off in range(0, len(a), step):
                assert np.array_equal(a[off:off + step], b[off:off + step]), f"{label}: JSONL bytes differ in [{off}, {off + step})"
        if flags & abi.RUN_LINKS:
            assert int(ro.n_links) == int(rg.n_links), f"{label}: link count"
            assert np.array_equal(_view(ro.link_off, n + 1, np.uint32), _view(rg.link_off, n + 1, np.uint32)), f"{label}: link_off"
            assert np.array_equal(_view(ro.links, int(ro.n_links) * 36, np.uint8), _view(rg.links, int(rg.n_links) * 36, np.uint8)), f"{label}: links"
        if flags & abi.RUN_FRONTIER:
            assert (int(ro.n_new), int(ro.frontier_size)) == (int(rg.n_new), int(rg.frontier_size)), f"{label}: frontier counters"
        return int(rg.jsonl_len), int(rg.n_links)
    finally:
        eng_mod.lib().tgi_result_release(e.h, rg.slot)


def test_config2_full_size_byte_parity():
    """bench --config 2: 10 M mixed Telegram messages, everything on, every byte compared"""
    n = max(1000, int(10_000_000 * _scale()))
    flags = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
    c = Corpus(n, seed=0x5EED0002, profile=2)
    e, o = Engine(frontier_capacity=1 << 23), Oracle()
    jl, nl = _compare_batch(e, o, c.batch, flags, "tg", "config 2")
    assert jl > 1500 * n and nl > 0
    assert np.array_equal(e.frontier_export(), o.frontier_export()), "frontier content / order"
    e.close()


def test_config3_full_size_links_and_frontier():
    """bench --config 3: 100 M link-bearing messages in three batches through ONE frontier; per-record links of
    every record and the final set (content and order) against the oracle"""
    n = max(3000, int(100_000_000 * _scale()))
    flags = abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
    e, o = Engine(frontier_capacity=1 << 25), Oracle()
    parts = [(n * i // 3, n * (i + 1) // 3) for i in range(3)]
    links = 0
    for k, (a, b) in enumerate(parts):
        c = Corpus(b - a, seed=0x5EED0003, first=a, profile=3)
        links += _compare_batch(e, o, c.batch, flags, "tg", f"config 3 part {k}")[1]
        c.close()
    fe, fo = e.frontier_export(), o.frontier_export()
    assert len(fe) == len(fo) and np.array_equal(fe, fo), "frontier content / order"
    assert links > len(fe) > 0
    e.close()


def test_config4_full_size_byte_parity():
    """bench --config 4: 50 M YouTube records streamed in batches; every JSONL byte, link and the frontier"""
    n = max(2000, int(50_000_000 * _scale()))
    per = min(n, 5_000_000)
    flags = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER
    e, o = Engine(frontier_capacity=1 << 23), Oracle()
    total = 0
    for a in range(0, n, per):
        c = YtCorpus(min(per, n - a), seed=0x5EED0004, first=a)
        total += _compare_batch(e, o, c.batch, flags, "yt", f"config 4 records [{a}, {a + c.batch.n})")[0]
        c.close()
    assert total > 3000 * n
    assert np.array_equal(e.frontier_export(), o.frontier_export()), "frontier content / order"
    e.close()


def test_staged_input_equals_plain_input():
    """inputs packed into tgi_acquire_staging memory give the same result as caller-owned arrays"""
    c = Corpus(50_000, profile=2)
    e = Engine()
    flags = abi.RUN_JSONL | abi.RUN_LINKS
    a = e.telegram(c.batch, flags)
    st = e.stage(c.batch)
    b = e.telegram(st, flags)
    e.unstage(st)
    st2 = e.stage(c.batch)  # the pool hands the block out again
    e.unstage(st2)
    assert np.array_equal(a.jsonl, b.jsonl) and np.array_equal(a.links, b.links) and np.array_equal(a.status, b.status)
    e.close()
