"""Combined-blob sink: groups the JSONL lines of one engine result into the blobs the reference's
chunk combiner would upload as combined_<ns>.jsonl (chunk/main.go:292-421), without one file per post.

The grouping rule is libtgingest's tgi_plan_chunks (a restatement of Chunker.processBatches); the bytes of a group are
a contiguous slice of the result's JSONL blob (minus lines dropped for exceeding the hard cap)."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import engine

TRIGGER_DEFAULT = 170 * 1024 * 1024  # the deployment's trigger / hard cap
HARD_CAP_DEFAULT = 200 * 1024 * 1024


def plan_chunks(line_off: np.ndarray, trigger: int = TRIGGER_DEFAULT, hard_cap: int = HARD_CAP_DEFAULT):
    """-> (groups [(begin, end)], dropped uint8[n])"""
    line_off = np.ascontiguousarray(line_off, dtype=np.uint64)
    n = len(line_off) - 1
    cap = max(n, 1)
    groups = np.zeros(2 * cap, np.uint64)
    dropped = np.zeros(max(n, 1), np.uint8)
    ng = C.c_uint64()
    rc = engine.lib().tgi_plan_chunks(line_off.ctypes.data, n, trigger, hard_cap, groups.ctypes.data, cap, C.byref(ng), dropped.ctypes.data)
    if rc:
        raise RuntimeError(f"tgi_plan_chunks: {rc}")
    g = groups[: 2 * ng.value].reshape(-1, 2)
    return [(int(a), int(b)) for a, b in g], dropped[:n]


def write_combined(jsonl: bytes | np.ndarray, line_off: np.ndarray, combine_dir: str, trigger: int = TRIGGER_DEFAULT,
                   hard_cap: int = HARD_CAP_DEFAULT, now_ns=time.time_ns) -> list[str]:
    """Writes one combined_<ns>.jsonl per group (chunk/main.go:378-380 naming); returns the paths."""
    buf = memoryview(jsonl)
    groups, dropped = plan_chunks(line_off, trigger, hard_cap)
    paths = []
    for a, b in groups:
        path = os.path.join(combine_dir, "combined_%d.jsonl" % now_ns())
        with open(path, "wb") as f:
            if not dropped[a:b].any():
                f.write(buf[int(line_off[a]): int(line_off[b])])
            else:
                for i in range(a, b):
                    if not dropped[i]:
                        f.write(buf[int(line_off[i]): int(line_off[i + 1])])
        paths.append(path)
    return paths


def plan_channel_appends(line_off: np.ndarray, recs: np.ndarray) -> np.ndarray:
    """runs of consecutive lines of one channel (tgi_plan_channel_appends) -> abi.APPEND_RUN[]"""
    from . import abi
    line_off = np.ascontiguousarray(line_off, dtype=np.uint64)
    n = len(line_off) - 1
    runs = np.zeros(max(n, 1), abi.APPEND_RUN)
    ng = C.c_uint64()
    base = recs.ctypes.data + recs.dtype.fields["chan_idx"][1] if n else None
    rc = engine.lib().tgi_plan_channel_appends(line_off.ctypes.data, base, recs.dtype.itemsize, n, runs.ctypes.data, len(runs), C.byref(ng))
    if rc:
        raise RuntimeError(f"tgi_plan_channel_appends: {rc}")
    return runs[: ng.value]


def append_posts(jsonl: bytes | np.ndarray, line_off: np.ndarray, recs: np.ndarray, channel_ids: list[str], base_path: str, crawl_id: str) -> int:
    """LocalStateManager.StorePost for a whole result (state/storageproviders.go:275-298): <base>/<crawl>/<channel>/posts/
    posts.jsonl gets the channel's lines, one append per run instead of one open / append / close per post.  Returns the
    number of appends."""
    buf = memoryview(jsonl)
    runs = plan_channel_appends(line_off, recs)
    for r in runs:
        d = os.path.join(base_path, crawl_id, channel_ids[int(r["chan_idx"])], "posts")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "posts.jsonl"), "ab") as f:
            f.write(buf[int(r["byte_begin"]): int(r["byte_end"])])
    return len(runs)
