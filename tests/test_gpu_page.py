"""Page-sized Telegram batches (crawl/runner.go:1110 hands ParseMessage 100 messages at a time): the one-launch path
(csrc/tg_page.cuh) against the oracle AND against the ordinary multi-launch pipeline, byte for byte."""
import os

import numpy as np
import pytest

from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
from distributed_crawler_b200.pack import pack_telegram
from helpers import ALL, TANDEM, assert_results_equal, msg, no_page
from oracle.pyoracle import Oracle
from test_gpu_parity import both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("profile", [1, 2, 3])
@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 100, 257, 1000, 2049, 4000])
def test_page_parity(n, profile):
    c = Corpus(n, profile=profile, first=777 + n)
    _, rg = both(c.batch, ALL)  # both(): oracle == page path == ordinary pipeline
    assert rg.gpu_launches == 1, "a page-sized batch must take the one-launch path"


@pytest.mark.parametrize("flags", [abi.RUN_JSONL, abi.RUN_LINKS, abi.RUN_LINKS | abi.RUN_FRONTIER, TANDEM,
                                   abi.RUN_JSONL | abi.RUN_FRONTIER, ALL | abi.RUN_FILTER])
def test_page_run_flags(flags):
    c = Corpus(300, profile=2, first=99)
    _, rg = both(c.batch, flags, tz_offset_sec=19800, crawl_label=b'p"<g>', min_post_date=1_720_000_000)
    assert rg.gpu_launches == 1


def test_pages_in_sequence_share_the_frontier():
    """Twenty pages, one big batch in between: every result and the final frontier equal the oracle's."""
    c = Corpus(2000, profile=2, first=5)
    big = Corpus(30000, profile=2, first=100000)
    o, e = Oracle(), Engine()
    for k in range(20):
        page = c.batch.slice(k * 100, (k + 1) * 100)
        ro, rg = o.telegram(page, ALL), e.telegram(page, ALL)
        assert rg.gpu_launches == 1
        assert_results_equal(ro, rg, ALL, f"page {k}")
        if k == 9:
            ro, rg = o.telegram(big.batch, ALL), e.telegram(big.batch, ALL)
            assert rg.gpu_launches > 1
            assert_results_equal(ro, rg, ALL, "big batch")
    assert np.array_equal(o.frontier_export(), e.frontier_export())
    e.close()


def test_page_speculative_read_grows():
    """Short pages shrink the estimate of the result size; a page of long lines then needs the second copy."""
    short = pack_telegram([msg("messageText", "a") for _ in range(200)])
    long_ = pack_telegram([msg("messageText", "line\n" * 3000 + " t.me/chan_%04d" % i) for i in range(200)])
    o, e = Oracle(), Engine()
    for b in (short, short, short, short, long_, short, long_):
        ro, rg = o.telegram(b, ALL), e.telegram(b, ALL)
        assert rg.gpu_launches == 1
        assert_results_equal(ro, rg, ALL)
    e.close()


def test_page_falls_back_when_the_arena_or_the_block_is_too_small():
    many = pack_telegram([msg("messageText", " ".join("t.me/c%03d_%05d" % (i, k) for k in range(150))) for i in range(40)])
    ro, rg = both(many, ALL)  # 6000 link candidates > the page arena: the ordinary pipeline answers
    assert rg.gpu_launches > 1 and rg.n_links == ro.links.shape[0]
    c = Corpus(500, profile=2, first=31)
    os.environ["TGI_PAGE_VAR_CAP"] = "65536"  # smaller than the page's JSONL
    try:
        o, e = Oracle(), Engine()
        ro, rg = o.telegram(c.batch, ALL), e.telegram(c.batch, ALL)
        assert rg.gpu_launches > 1
        assert_results_equal(ro, rg, ALL)
        assert np.array_equal(o.frontier_export(), e.frontier_export())  # nothing was committed twice
        e.close()
    finally:
        del os.environ["TGI_PAGE_VAR_CAP"]


def test_page_resident_rerun_and_device_read():
    c = Corpus(400, profile=3, first=8)
    o, e = Oracle(), Engine()
    e.telegram_upload(0, c.batch)
    for flags in (abi.RUN_JSONL, ALL, TANDEM):
        o2 = Oracle()
        ro = o2.telegram(c.batch, flags)
        e.frontier_clear()
        rg = e.telegram_run_resident(0, flags, copy=True)
        assert rg.gpu_launches == 1
        assert_results_equal(ro, rg, flags)
        if flags & abi.RUN_JSONL:
            assert e.read_jsonl(0, 0, rg.jsonl_len) == ro.jsonl.tobytes()
        e.release(0)
    e.close()


def test_pages_on_three_slots():
    c = Corpus(900, profile=2, first=3)
    o, e = Oracle(), Engine()
    pages = [c.batch.slice(k * 300, (k + 1) * 300) for k in range(3)]
    for k, p in enumerate(pages):
        e.telegram_submit(k, p, ALL)
    for k, p in enumerate(pages):
        rg = e.telegram_wait(k, copy=True)
        ro = o.telegram(p, ALL)
        assert rg.gpu_launches == 1
        # the frontier phases run in submission order: the same split of "new" keys as one thread going page by page
        assert_results_equal(ro, rg, ALL, f"slot {k}")
        e.release(k)
    assert np.array_equal(o.frontier_export(), e.frontier_export())
    e.close()


def test_page_path_is_cheaper_than_the_pipeline():
    """Not a benchmark: the one-launch call must not be slower than the 19-launch one on the same page."""
    import time
    c = Corpus(100, profile=2, first=1)
    e = Engine()

    def best():
        ts = []
        for _ in range(30):
            e.frontier_clear()
            t = time.perf_counter()
            e.telegram(c.batch, ALL, copy=False)
            ts.append(time.perf_counter() - t)
        return min(ts)

    page = best()
    with no_page():
        pipe = best()
    e.close()
    assert page < pipe, (page, pipe)


# ---- YouTube pages (the Data API returns 50 videos per page: crawler/youtube/youtube_crawler.go:353-427) -------------------
import test_gpu_youtube as ytt
from yt_corpus import make_youtube, make_youtube_config4


@pytest.mark.parametrize("n", [1, 7, 50, 333, 2000])
@pytest.mark.parametrize("make", [make_youtube, make_youtube_config4])
def test_youtube_page_parity(n, make):
    batch, _, _ = make(n, seed=100 + n)
    _, rg = ytt.both(batch)  # oracle == page path == bulk pipeline
    assert rg.gpu_launches == 1


def test_youtube_page_flags_and_sequence():
    batch, _, _ = make_youtube(600, seed=5)
    for flags in (abi.RUN_JSONL, abi.RUN_LINKS, abi.RUN_LINKS | abi.RUN_FRONTIER):
        _, rg = ytt.both(batch, flags, tz_offset_sec=-18000, crawl_label=b"yt<page>")
        assert rg.gpu_launches == 1
    o, e = Oracle(), Engine()
    for k in range(8):  # pages of 50 sharing the snowball frontier
        page, _, _ = make_youtube(50, seed=900 + k % 3)  # every third page repeats: nothing new the second time
        ro, rg = o.youtube(page, ytt.ALL), e.youtube(page, ytt.ALL)
        assert rg.gpu_launches == 1
        assert_results_equal(ro, rg, ytt.ALL, f"page {k}")
    assert np.array_equal(o.frontier_export(), e.frontier_export())
    e.close()


def test_youtube_page_falls_back_when_the_block_is_too_small():
    batch, _, _ = make_youtube(400, seed=8)
    os.environ["TGI_PAGE_VAR_CAP"] = "65536"
    try:
        o, e = Oracle(), Engine()
        ro, rg = o.youtube(batch, ytt.ALL), e.youtube(batch, ytt.ALL)
        assert rg.gpu_launches > 1
        assert_results_equal(ro, rg, ytt.ALL)
        assert np.array_equal(o.frontier_export(), e.frontier_export())
        e.close()
    finally:
        del os.environ["TGI_PAGE_VAR_CAP"]


def test_page_hand_off_rows():
    """f3 on the page path: exclusion sets, TGI_RUN_SKIP_INVALID and the pending_edges rows of a page equal the oracle's."""
    from distributed_crawler_b200.engine import names_to_keys32
    from test_handoff import NOW, _sets
    c = Corpus(3000, profile=3, nthreads=2)
    plain = Oracle().telegram(c.batch, abi.RUN_LINKS)
    inv, stamps, disc = _sets(plain.links)
    o, e = Oracle(), Engine()
    for x in (o, e):
        x.set_add(abi.SET_INVALID, names_to_keys32(inv), np.array(stamps, np.int64))
        x.set_add(abi.SET_DISCOVERED, names_to_keys32(disc))
        x.set_now(NOW)
    T = TANDEM | abi.RUN_SKIP_INVALID
    total = 0
    for k in range(6):
        page = c.batch.slice(k * 500, (k + 1) * 500)
        ro = o.telegram(page, T)
        e.telegram_submit(2, page, T)
        rg = e.telegram_wait(2, copy=True)
        assert rg.gpu_launches == 1
        assert np.array_equal(ro.links, rg.links) and ro.n_new == rg.n_new
        rows_o, rows_g = o.pending_edges(NOW), e.pending_edges(2, NOW)
        assert np.array_equal(rows_o, rows_g) and len(rows_g) == rg.n_new
        total += len(rows_g)
        e.release(2)
    assert total > 0
    e.close()


def test_youtube_page_pending_edges_equal_the_bulk_pipelines():
    """tgi_pending_edges after a YouTube batch reads the resident descriptor: same rows from the page path and the bulk path."""
    batch, _, _ = make_youtube(500, seed=12)
    rows = []
    for bulk in (False, True):
        e = Engine()
        if bulk:
            os.environ["TGI_NO_PAGE"] = "1"
        try:
            e.youtube_submit(1, batch, ytt.ALL)
            r = e.youtube_wait(1, copy=True)
        finally:
            os.environ.pop("TGI_NO_PAGE", None)
        assert (r.gpu_launches == 1) == (not bulk)
        got = e.pending_edges(1, 1_760_000_000)
        assert len(got) == r.n_new and r.n_new > 0
        rows.append(got.copy())
        e.release(1)
        e.close()
    assert np.array_equal(rows[0], rows[1])


@pytest.mark.parametrize("bulk", [False, True])
def test_max_out_bytes_rejects_the_batch_without_touching_the_frontier(bulk):
    """tgi_config.max_out_bytes: a batch whose JSONL is over the limit is TGI_E_CAPACITY and leaves the dedup set alone."""
    from distributed_crawler_b200.engine import EngineError
    c = Corpus(300, profile=2, first=4)
    e = Engine(max_out_bytes=10_000)
    if bulk:
        os.environ["TGI_NO_PAGE"] = "1"
    try:
        with pytest.raises(EngineError) as ei:
            e.telegram(c.batch, ALL)
        assert ei.value.code == abi.E_CAPACITY
    finally:
        os.environ.pop("TGI_NO_PAGE", None)
    assert e.frontier_size() == 0
    ro, rg = Oracle().telegram(c.batch, abi.RUN_LINKS | abi.RUN_FRONTIER), e.telegram(c.batch, abi.RUN_LINKS | abi.RUN_FRONTIER)
    assert ro.n_new == rg.n_new > 0  # links-only batches carry no JSONL: not limited
    e.close()
