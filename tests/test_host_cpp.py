"""The C++ host mirror (distributed_crawler_b200/host/tgingest.hpp): the same messages packed by the C++
Batch builder and by pack.py must give identical arrays (CPU), and identical results through the engine (GPU)."""
import os
import subprocess

import numpy as np
import pytest

from distributed_crawler_b200 import abi
from distributed_crawler_b200.pack import Channel, Comment, FormattedText, Message, TextEntity, pack_telegram

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "distributed_crawler_b200", "host")


def fixture():
    chans = [Channel("Test Channel", "testchannel", "testchannel", 1200, 34, 56789), Channel('Приватный "канал"', "private_chan", "", 0, 0, 0)]
    msgs = [
        Message(id=5 << 20, chat_id=-1001234567890, date=1700000000, view_count=1234, share_count=7,
                text=FormattedText("Join @durov_channel and t.me/some_channel now\nsecond line <b>", [TextEntity(5, 14, "mention")]),
                reactions=[("👍", 12), ("❤️", 3)]),
        Message(id=6 << 20, chat_id=-1001234567890, date=1700000100, content_type="messageVideo", media_album_id=99, media="BAACAgIAAxkBAAIB",
                text=FormattedText("caption with a link", [TextEntity(15, 4, "text_url", "https://t.me/linked_channel/42")]),
                comments=[Comment("first!", [("🔥", 2)], 10, 1, "someone"), Comment("no reactions", None, 0, 0, "unknown")]),
        Message(id=7 << 20, chat_id=-1009876543210, date=1600000000, content_type="messagePoll", alt="What do you think?", comments=None, channel=1),
        Message(id=8 << 20, chat_id=-1009876543210, date=1650000000, content_type="messageDice", alt="messageDice", channel=1, panics=True),
    ]
    return pack_telegram(msgs, chans)


def fnv(a: np.ndarray) -> str:
    h = 1469598103934665603
    for b in a.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def demo(*args):
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return subprocess.run([os.path.join(HOST, "host_demo"), *args], capture_output=True, timeout=300)


def test_cpp_batch_builder_packs_like_pack_py(engine_lib):
    p = demo("--pack")
    assert p.returncode == 0, p.stderr
    got = dict(l.split() for l in p.stdout.decode().splitlines())
    b = fixture()
    assert int(got.pop("n")) == b.n
    for k in ("recs", "strs", "ent_off", "ents", "react_off", "reacts", "comment_off", "comments", "aux", "chans", "chan_strs"):
        assert got[k] == fnv(getattr(b, k)), k


@pytest.mark.gpu
def test_cpp_message_processor_matches_python_engine():
    from distributed_crawler_b200.engine import Engine
    p = demo("--run")
    assert p.returncode == 0, p.stderr
    e = Engine(crawl_label=b'demo "label"', tz_offset_sec=3600, created_at_sec=1750000000, created_at_nsec=0, capture_sec=1750000001, capture_nsec=500)
    r = e.telegram(fixture(), abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF)
    want = b""
    for i in range(r.n):
        names = [bytes(l["name"][: l["len"]]).decode() for l in r.links[r.link_off[i]: r.link_off[i + 1]]]
        want += ("status %d links" % r.status[i]).encode() + "".join(" " + n for n in names).encode() + b"\n" + r.line(i)
    assert p.stdout == want
    assert list(r.status) == [abi.ST_EMITTED, abi.ST_EMITTED, abi.ST_EMITTED, abi.ST_FAILED]
    e.close()
