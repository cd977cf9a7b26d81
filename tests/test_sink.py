"""tgi_plan_chunks / sink.write_combined against a restatement of Chunker.processBatches
(chunk/main.go:292-345) that works on a list of per-post files, as the reference does."""
import random

import numpy as np

from distributed_crawler_b200 import sink


def process_batches(sizes, trigger, hard_cap):
    """chunk/main.go:292-345 on FileEntry{Path: index, Size}: returns the batches (lists of file indices)."""
    out, files, size = [], [], 0

    def flush():
        nonlocal files, size
        if files:
            out.append(files)
            files, size = [], 0

    for i, sz in enumerate(sizes):
        if sz > hard_cap:          # :316-322 deleted
            continue
        if size > 0 and size + sz > hard_cap:  # :324-327
            flush()
        files.append(i)
        size += sz
        if size >= trigger:        # :334-337
            flush()
    flush()                        # :339-343
    return out


def test_plan_chunks_matches_process_batches(engine_lib):
    rnd = random.Random(4)
    for trial in range(200):
        n = rnd.randrange(0, 120)
        trigger = rnd.randrange(1, 4000)
        hard_cap = trigger + rnd.randrange(0, 1500)
        lens = [rnd.choice([0, 0, rnd.randrange(1, 900), rnd.randrange(1, 900), rnd.randrange(1, 6000)]) for _ in range(n)]
        line_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        groups, dropped = sink.plan_chunks(line_off, trigger, hard_cap)
        # the reference only ever sees posts that produced a line: one file per non-empty line
        file_idx = [i for i, l in enumerate(lens) if l > 0]
        want = [[file_idx[k] for k in batch] for batch in process_batches([lens[i] for i in file_idx], trigger, hard_cap)]
        got = [[i for i in range(a, b) if lens[i] > 0 and not dropped[i]] for a, b in groups]
        assert got == want, (trial, lens, trigger, hard_cap)
        assert [i for i in range(n) if dropped[i]] == [i for i in file_idx if lens[i] > hard_cap]
        assert all(groups[k][1] <= groups[k + 1][0] for k in range(len(groups) - 1))


def test_write_combined_concatenates_lines(engine_lib, tmp_path):
    lines = [b'{"a":%d}\n' % i * (1 + i % 3) for i in range(50)]
    lines[7] = b""
    lines[20] = b"x" * 500 + b"\n"  # over the hard cap: dropped
    blob = b"".join(lines)
    line_off = np.concatenate([[0], np.cumsum([len(l) for l in lines])]).astype(np.uint64)
    t = iter(range(1000, 2000))
    paths = sink.write_combined(blob, line_off, str(tmp_path), trigger=100, hard_cap=130, now_ns=lambda: next(t))
    assert [p.rsplit("/", 1)[1] for p in paths[:2]] == ["combined_1000.jsonl", "combined_1001.jsonl"]
    got = b"".join(open(p, "rb").read() for p in paths)
    assert got == b"".join(l for i, l in enumerate(lines) if i != 20)
    assert all(len(open(p, "rb").read()) <= 130 for p in paths)


def test_channel_appends_equal_per_post_appends(engine_lib, tmp_path):
    """tgi_plan_channel_appends / sink.append_posts against LocalStateManager.StorePost called once per post
    (state/storageproviders.go:39-53,275-298): identical posts.jsonl per channel, far fewer open/append/close cycles."""
    import os
    from distributed_crawler_b200 import abi
    rnd = random.Random(9)
    n, n_chans = 400, 7
    recs = np.zeros(n, abi.TG_REC)
    chan, lines = 0, []
    for i in range(n):
        if rnd.random() < 0.1:
            chan = rnd.randrange(n_chans)  # pages of several channels in one batch; channels come back later
        recs["chan_idx"][i] = chan
        lines.append(b"" if rnd.random() < 0.15 else b'{"i":%d,"c":%d}\n' % (i, chan))
    blob = b"".join(lines)
    line_off = np.concatenate([[0], np.cumsum([len(l) for l in lines])]).astype(np.uint64)
    names = ["chan%d" % c for c in range(n_chans)]
    ref = {}
    for i, l in enumerate(lines):  # StorePost per post: append to the channel's file
        if l:
            ref.setdefault(names[int(recs["chan_idx"][i])], []).append(l)
    appends = sink.append_posts(blob, line_off, recs, names, str(tmp_path), "crawl1")
    for name, ls in ref.items():
        assert open(os.path.join(tmp_path, "crawl1", name, "posts", "posts.jsonl"), "rb").read() == b"".join(ls)
    runs = sink.plan_channel_appends(line_off, recs)
    assert appends == len(runs) < sum(1 for l in lines if l) / 3
    assert int(runs["n_lines"].sum()) == sum(1 for l in lines if l)
    assert sink.plan_channel_appends(np.zeros(1, np.uint64), recs[:0]).size == 0


def test_reference_process_batches_vectors(engine_lib, vectors):
    """The reference's own known-answer tests of the batching rule (chunk/main_test.go TestProcessBatches_*), through
    tgi_plan_chunks and through the restatement above."""
    for v in vectors["chunk_batches"]:
        sizes, where = v["sizes"], f'{v["name"]} ({v["go_file"]}:{v["go_line"]})'
        line_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        groups, dropped = sink.plan_chunks(line_off, v["trigger"], v["hard_cap"])
        got = [[i for i in range(a, b) if not dropped[i]] for a, b in groups]
        assert got == process_batches(sizes, v["trigger"], v["hard_cap"]), where
        ex = v["expect"]
        if "batches" in ex:
            assert got == ex["batches"], where
        if "min_batches" in ex:
            assert len(got) >= ex["min_batches"], where
        if "first_batches" in ex:
            assert got[: len(ex["first_batches"])] == ex["first_batches"], where
        if "files_in_batches" in ex:
            assert sum(len(b) for b in got) == ex["files_in_batches"], where
        if "dropped" in ex:
            assert [i for i in range(len(sizes)) if dropped[i]] == ex["dropped"], where
        if "total_size" in ex:  # Chunker.totalUploadSize after the run
            assert sum(sizes[i] for b in got for i in b) == ex["total_size"], where
