"""Builders that mirror the helpers of the reference's own tests
(telegramhelper/channel_links_test.go:34-60: msgText, msgPhoto, ...)."""
from __future__ import annotations

import contextlib
import os

import numpy as np

from distributed_crawler_b200 import abi
from distributed_crawler_b200.pack import Channel, FormattedText, Message, TextEntity, pack_telegram


def ft(text, entities=()):
    if text is None:
        return None
    return FormattedText(text, [TextEntity(o, l, t, u) for (o, l, t, u) in entities])


def msg(content_type="messageText", text=None, entities=(), **kw):
    return Message(content_type=content_type, text=ft(text, entities), **kw)


def vector_message(v):
    return msg(v["content_type"], v["text"], [tuple(e) for e in v["entities"]])


def names(result, i=0):
    return sorted(n.decode() for n, _ in result.record_links(i))


def assert_results_equal(ro, rg, flags, label=""):
    assert np.array_equal(ro.status, rg.status), f"{label}: status differs"
    if flags & abi.RUN_JSONL:
        if not np.array_equal(ro.line_off, rg.line_off) or not np.array_equal(ro.jsonl, rg.jsonl):
            for i in range(ro.n):
                a, c = ro.line(i), rg.line(i)
                if a != c:
                    j = next((k for k in range(min(len(a), len(c))) if a[k] != c[k]), min(len(a), len(c)))
                    raise AssertionError(f"{label}: line {i} differs at byte {j}: "
                                         f"oracle={a[max(0, j - 60):j + 60]!r} gpu={c[max(0, j - 60):j + 60]!r}")
            raise AssertionError(f"{label}: offsets differ")
    if flags & abi.RUN_LINKS:
        assert np.array_equal(ro.link_off, rg.link_off), f"{label}: link_off differs"
        assert np.array_equal(ro.links, rg.links), f"{label}: links differ"
    if flags & abi.RUN_FRONTIER:
        assert ro.n_new == rg.n_new, f"{label}: n_new {ro.n_new} != {rg.n_new}"
        assert ro.frontier_size == rg.frontier_size, f"{label}: frontier size"


ALL = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
TANDEM = abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_FILTER | abi.RUN_SKIP_SELF


@contextlib.contextmanager
def no_page():
    """The ordinary multi-launch pipeline also for page-sized batches (the library reads TGI_NO_PAGE per call)."""
    os.environ["TGI_NO_PAGE"] = "1"
    try:
        yield
    finally:
        del os.environ["TGI_NO_PAGE"]
