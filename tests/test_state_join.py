"""the message-status join.  CPU: the oracle's join against a literal restatement of the reference's
map-based functions (crawl/runner.go:1572-1697).  GPU: tgi_key_join against the oracle."""
import random

import numpy as np
import pytest

from distributed_crawler_b200 import state_join
from oracle.pyoracle import Oracle


def go_resample_marker(messages, discovered):  # crawl/runner.go:1572-1635 (messages: [(chat, id, status)])
    dmap = {"%d_%d" % (c, m): True for c, m in discovered}
    out = []
    for c, m, st in messages:
        if st == "fetched":
            out.append(st)
        elif dmap.get("%d_%d" % (c, m)):
            out.append("resample")
        else:
            out.append("deleted")
    return out


def go_add_new_messages(discovered, existing):  # crawl/runner.go:1650-1697 -> indices of discovered that are appended
    emap = {"%d_%d" % (c, m): True for c, m in existing}
    return [i for i, (c, m) in enumerate(discovered) if not emap.get("%d_%d" % (c, m))]


def lists(rnd, n, m):
    chats = [-1000000000000 - rnd.randrange(5) for _ in range(3)] + [0, 7]
    pool = [(rnd.choice(chats), rnd.randrange(-3, 60) << 20) for _ in range(80)]
    a = [rnd.choice(pool) for _ in range(n)]
    b = [rnd.choice(pool) for _ in range(m)]
    return a, b


def check(join):
    rnd = random.Random(11)
    for trial in range(60):
        existing, discovered = lists(rnd, rnd.randrange(0, 70), rnd.randrange(0, 70))
        status = [rnd.choice(["fetched", "unfetched", "failed", "resample"]) for _ in existing]
        ek, dk = np.array(existing, np.int64).reshape(-1, 2), np.array(discovered, np.int64).reshape(-1, 2)
        assert state_join.resample_marker(join, ek, status, dk) == go_resample_marker([(c, m, s) for (c, m), s in zip(existing, status)], discovered)
        assert list(state_join.add_new_messages(join, dk, ek)) == go_add_new_messages(discovered, existing)
        want = [next((i for i, k in enumerate(discovered) if k == key), -1) for key in existing]  # :1171-1176 linear search
        assert list(state_join.find_fetched(join, dk, ek)) == want


def test_oracle_join_matches_the_reference_maps():
    check(Oracle.key_join)


@pytest.mark.gpu
def test_gpu_join_matches_oracle():
    from distributed_crawler_b200.engine import Engine
    e = Engine()
    check(e.key_join)
    rnd = np.random.default_rng(3)
    a = rnd.integers(-50000, 50000, size=(300000, 2))
    b = rnd.integers(-50000, 50000, size=(200000, 2))
    assert np.array_equal(e.key_join(a, b), Oracle.key_join(a, b))
    e.close()


def test_update_messages_batch_equals_sequential_update_message():
    """state_join.update_messages_batch against BaseStateManager.UpdateMessage (state/base.go:182-215) called once per
    update: first match wins, unknown keys append, later updates hit what was appended."""
    import random
    from distributed_crawler_b200.state_join import update_messages_batch
    rnd = random.Random(3)
    for trial in range(50):
        page = [(rnd.randrange(3), rnd.randrange(12)) for _ in range(rnd.randrange(0, 25))]  # duplicates on purpose
        status = [rnd.choice(["unfetched", "fetched"]) for _ in page]
        ups = [((rnd.randrange(3), rnd.randrange(16)), rnd.choice(["fetched", "failed", "deleted"])) for _ in range(rnd.randrange(0, 40))]
        want_keys, want_status = list(page), list(status)
        for k, st in ups:  # the reference, one call per update
            for i, pk in enumerate(want_keys):
                if pk == k:
                    want_status[i] = st
                    break
            else:
                want_keys.append(k)
                want_status.append(st)
        keys, got = update_messages_batch(Oracle.key_join, np.array(page, np.int64).reshape(-1, 2), status,
                                          np.array([k for k, _ in ups], np.int64).reshape(-1, 2), [s for _, s in ups])
        assert [tuple(map(int, k)) for k in keys] == want_keys and got == want_status, trial
