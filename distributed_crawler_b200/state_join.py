"""Message-status bookkeeping of crawl/runner.go on top of ONE join primitive.

`join(a_keys, b_keys)` returns, for every (chat_id, message_id) of b, the index of the first equal element of a or -1:
`Engine.key_join` on the GPU (tgi_key_join), `Oracle.key_join` on the CPU.  The functions below restate what the
reference does with string-keyed maps / linear scans, in terms of that primitive."""
from __future__ import annotations

import numpy as np

FETCHED, RESAMPLE, DELETED = "fetched", "resample", "deleted"


def resample_marker(join, existing_keys, existing_status: list[str], discovered_keys) -> list[str]:
    """resampleMarker (crawl/runner.go:1572-1635): fetched stays; found among the discovered -> "resample"; else "deleted"."""
    idx = join(discovered_keys, existing_keys)
    return [s if s == FETCHED else (RESAMPLE if idx[i] >= 0 else DELETED) for i, s in enumerate(existing_status)]


def add_new_messages(join, discovered_keys, existing_keys) -> np.ndarray:
    """addNewMessages (crawl/runner.go:1650-1697): indices of the discovered messages that are not in owner.Messages, in
    order (a message discovered twice is appended twice: the reference does not update its map while it appends)."""
    idx = join(existing_keys, discovered_keys)
    return np.nonzero(idx < 0)[0]


def find_fetched(join, fetched_keys, state_keys) -> np.ndarray:
    """the per-message search crawl/runner.go:1171-1176: for every state message the first fetched message with its id"""
    return join(fetched_keys, state_keys)
