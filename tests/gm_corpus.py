"""Seeded generic client.Message fixtures for the oracle and GPU tests."""
import random

from distributed_crawler_b200.pack import GenericMessage, pack_generic

_WORDS = ["hello", "мир", "世界", "t.me/somechan", "line\nbreak", "tab\there", 'quote"d', "back\\slash", "<b>&amp;</b>",
          "emoji😀", " sep", "ok", "https://example.org/a?b=c", "ctl\x01\x1f", "é", ""]
_KEYS = ["👍", "❤️", "🔥", "a", "zz", 'k"q', "long-reaction-key-number-one", "😀", "\x7f", "<"]


def make_generic(n: int, seed: int = 5):
    rnd = random.Random(seed)
    msgs = []
    for i in range(n):
        text = " ".join(rnd.choice(_WORDS) for _ in range(rnd.randrange(0, 40))).encode()
        if rnd.random() < 0.05:
            text += bytes(rnd.randrange(0x80, 0x100) for _ in range(rnd.randrange(1, 5)))  # invalid UTF-8
        k = rnd.randrange(0, 5) if rnd.random() < 0.7 else rnd.randrange(0, 12)
        reactions = [(rnd.choice(_KEYS), rnd.randrange(-3, 10 ** rnd.randrange(1, 12))) for _ in range(k)]  # duplicates happen
        msgs.append(GenericMessage(
            id=str(rnd.randrange(1, 1 << 50)), channel_id=rnd.choice(["chan_one", "Ünïcode", 'we"ird', ""]), text=text,
            sender_name=rnd.choice(["", "", "someone", "q\"x"]), ts_sec=rnd.randrange(1_600_000_000, 1_800_000_000),
            ts_nsec=0, views=rnd.choice([0, 1, 12345, rnd.randrange(0, 1 << 40), -7]), reactions=reactions))
    return pack_generic(msgs), msgs
