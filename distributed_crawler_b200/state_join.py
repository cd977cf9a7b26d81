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


def update_messages_batch(join, page_keys, page_status: list[str], update_keys, update_status: list[str]):
    """BaseStateManager.UpdateMessage (state/base.go:182-215) for a whole batch of (chat_id, message_id, status) updates:
    the reference scans page.Messages linearly for EVERY update (O(messages x updates)) and the local manager rewrites
    state.json after each one (state/storageproviders.go:431-440).  Here: one join, then one pass in update order —
    an update whose key is in the page sets the status of the FIRST message with that key, an unknown key appends a new
    message (and later updates of that key hit the appended one), exactly like the sequential calls; the caller persists
    the page once.  Returns (keys int64[m,2], status list[str]) of the updated page."""
    page_keys = np.ascontiguousarray(page_keys, dtype=np.int64).reshape(-1, 2)
    update_keys = np.ascontiguousarray(update_keys, dtype=np.int64).reshape(-1, 2)
    idx = join(page_keys, update_keys)
    status = list(page_status)
    appended: dict[tuple[int, int], int] = {}
    extra = []
    for j, st in enumerate(update_status):
        i = int(idx[j])
        if i < 0:
            k = (int(update_keys[j, 0]), int(update_keys[j, 1]))
            i = appended.get(k, -1)
            if i < 0:
                i = appended[k] = len(status)
                status.append(st)
                extra.append(k)
                continue
        status[i] = st
    keys = np.concatenate([page_keys, np.array(extra, np.int64).reshape(-1, 2)]) if extra else page_keys
    return keys, status
