"""GPU parity for the generic client.Message -> sparse Post path (SURVEY §8 a12,
crawler/telegram/telegram_crawler.go:179-262): CUDA through the C ABI vs the CPU oracle, byte equality."""
import pytest

from distributed_crawler_b200 import abi
from distributed_crawler_b200.engine import Engine
from distributed_crawler_b200.pack import GenericMessage, pack_generic
from gm_corpus import make_generic
from helpers import assert_results_equal
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


def both(batch, flags=abi.RUN_JSONL, **cfg):
    o, e = Oracle(**cfg), Engine(**cfg)
    ro, rg = o.generic(batch, flags), e.generic(batch, flags)
    assert_results_equal(ro, rg, flags)
    assert rg.gpu_launches > 0
    e.close()
    return ro, rg


@pytest.mark.parametrize("n", [1, 7, 4000])
def test_generic_corpus_parity(n):
    batch, _ = make_generic(n, seed=3 + n)
    both(batch)
    both(batch, abi.RUN_JSONL | abi.RUN_LINKS, tz_offset_sec=-12600, created_at_nsec=123_000_000, capture_nsec=5)


def test_generic_edge_cases():
    both(pack_generic([]))
    msgs = [GenericMessage(),  # everything empty: nil reactions -> null
            GenericMessage(id="1", channel_id="c", text="x" * 5000, ts_sec=253402300800),  # year 10000: Marshal error -> no line
            GenericMessage(id="2", channel_id="c", text="ok", ts_sec=-62135596800, views=-(1 << 63),
                           reactions=[("b", 1), ("a", 2), ("b", 3), ("", 4), ("a\x00", 5)])]
    ro, rg = both(pack_generic(msgs))
    assert list(rg.status) == [abi.ST_EMITTED, abi.ST_NOLINE, abi.ST_EMITTED]
    assert b'"reactions":{"":4,"a":2,"a\\u0000":5,"b":3}' in rg.line(2)
    assert b'"reactions":null' in rg.line(0) and b'"platform_name":"telegram"' in rg.line(0)
    ro2, rg2 = both(pack_generic(msgs), abi.RUN_LINKS)  # the status does not depend on TGI_RUN_JSONL
    assert list(rg2.status) == list(rg.status)
