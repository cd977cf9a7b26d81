"""GPU parity for the YouTube path (BASELINE config 4 shape): convertVideoToPost + JSONL, extractURLs,
extractChannelIDsFromText, frontier — CUDA through the C ABI vs the CPU oracle, byte equality."""
import numpy as np
import pytest

from distributed_crawler_b200 import abi
from distributed_crawler_b200.engine import Engine
from distributed_crawler_b200.pack import YouTubeChannel, YouTubeVideo, pack_youtube
from helpers import assert_results_equal, no_page
from oracle.pyoracle import Oracle
from yt_corpus import make_youtube

pytestmark = pytest.mark.gpu
ALL = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER


def both(batch, flags=ALL, **cfg):
    o, e = Oracle(**cfg), Engine(**cfg)
    ro, rg = o.youtube(batch, flags), e.youtube(batch, flags)
    assert_results_equal(ro, rg, flags)
    if flags & abi.RUN_FRONTIER:
        assert np.array_equal(o.frontier_export(), e.frontier_export())
    assert rg.gpu_launches > 0
    if rg.gpu_launches == 1 and batch.n:  # a page-sized batch took the one-launch path: the bulk pipeline as well
        with no_page():
            e2 = Engine(**cfg)
            r2 = e2.youtube(batch, flags)
            assert r2.gpu_launches > 1
            assert_results_equal(ro, r2, flags, "bulk pipeline")
            if flags & abi.RUN_FRONTIER:
                assert np.array_equal(o.frontier_export(), e2.frontier_export())
            e2.close()
    e.close()
    return ro, rg


@pytest.mark.parametrize("n", [1, 5, 3000, 20000])
def test_youtube_corpus_parity(n):
    batch, _, _ = make_youtube(n, seed=11 + n)
    both(batch)


def test_youtube_configs_and_edge_cases():
    batch, _, _ = make_youtube(2000, seed=3)
    both(batch, ALL, tz_offset_sec=19800, crawl_label=b'yt"<lbl>', created_at_nsec=987_000_000, capture_nsec=0)
    both(batch, abi.RUN_JSONL)
    both(batch, abi.RUN_LINKS)
    both(pack_youtube([]))
    vids = [YouTubeVideo(id="", title="", description="", duration="", thumbnails={}),
            YouTubeVideo(id='a"b', title="x" * 300 + "é" * 40, description="http://a.b/c, (https://d.e/f). https://d.e/f 'http://g'", duration="PT1M",
                         thumbnails={"maxres": "m", "default": "d", "high": ""}, view_count=-5, like_count=-1),
            YouTubeVideo(description="see youtube.com/channel/UCabc-_123 and https://youtube.com/@some.handle-1/x youtube.com/@ youtube.com/channel/",
                         title="  \xe9", channel=1)]
    ro, rg = both(pack_youtube(vids, [YouTubeChannel(id="UCx", title="T", cached=True), YouTubeChannel(id="@h", cached=False)]))
    assert [bytes(l["name"][: l["len"]]) for l in rg.links] == [b"UCabc-_123", b"@some.handle-1"]


def test_youtube_config4_shape_parity():
    """BASELINE config 4 shape: almost every record takes the lane writer (nothing to escape)."""
    from yt_corpus import make_youtube_config4
    batch, _, _ = make_youtube_config4(3000, seed=77)
    ro, rg = both(batch)
    assert (rg.status == abi.ST_EMITTED).all()
