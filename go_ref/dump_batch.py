#!/usr/bin/env python3
"""Writes the seeded corpora used by tests/test_go_fixtures.py as raw packed-batch files for go_ref/main.go.

File = magic "TGB1" | "YTB1", u32 version, the tgi_config scalars, then every array of the batch descriptor in the
order of include/tgingest.h as  u64 byte length + bytes (padded to 8).  tests/test_go_fixtures.py reads the same
files back, so the Go harness, the oracle and the CUDA path see bit-identical inputs."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CFG = dict(tz_offset_sec=0, created_at_sec=1_750_000_000, created_at_nsec=0, capture_sec=1_750_000_000, capture_nsec=123_456_789,
           crawl_label=b"")
CASES_TG = [("tg_text_10k", 10_000, 1, 0x5EED0001), ("tg_mixed_20k", 20_000, 2, 0x5EED0002), ("tg_links_20k", 20_000, 3, 0x5EED0003)]
CASES_YT = [("yt_config4_5k", 5_000, 0x5EED0004)]


def write_batch(path, magic, batch, cfg=CFG):
    with open(path, "wb") as f:
        f.write(magic + struct.pack("<IiqiqiI", 1, cfg["tz_offset_sec"], cfg["created_at_sec"], cfg["created_at_nsec"],
                                    cfg["capture_sec"], cfg["capture_nsec"], len(cfg["crawl_label"])))
        f.write(cfg["crawl_label"].ljust((len(cfg["crawl_label"]) + 7) & ~7, b"\0"))
        for k in batch.FIELDS:
            a = np.ascontiguousarray(getattr(batch, k))
            raw = a.tobytes()
            f.write(struct.pack("<Q", len(raw)) + raw.ljust((len(raw) + 7) & ~7, b"\0"))


def read_batch(path):
    """-> (kind, batch, cfg)"""
    from distributed_crawler_b200 import abi
    from distributed_crawler_b200.pack import TgBatch, YtBatch
    raw = open(path, "rb").read()
    magic = raw[:4]
    ver, tz, cs, cn, ps, pn, ll = struct.unpack_from("<IiqiqiI", raw, 4)
    o = 4 + struct.calcsize("<IiqiqiI")
    label = raw[o:o + ll]
    o += (ll + 7) & ~7
    cls, dts = (TgBatch, dict(recs=abi.TG_REC, strs=np.uint8, ent_off=np.uint32, ents=abi.ENTITY, react_off=np.uint32,
                              reacts=abi.REACTION, comment_off=np.uint32, comments=abi.COMMENT, aux=np.uint8,
                              chans=abi.TG_CHAN, chan_strs=np.uint8)) if magic == b"TGB1" else \
               (YtBatch, dict(recs=abi.YT_REC, strs=np.uint8, chans=abi.YT_CHAN, chan_strs=np.uint8))
    arrays = {}
    for k in cls.FIELDS:
        (n,) = struct.unpack_from("<Q", raw, o)
        o += 8
        pad = np.zeros(n + 16, np.uint8)  # 16 readable bytes behind every array, as the packers guarantee
        pad[:n] = np.frombuffer(raw, np.uint8, n, o)
        arrays[k] = pad[:n].view(dts[k])
        o += (n + 7) & ~7
    cfg = dict(tz_offset_sec=tz, created_at_sec=cs, created_at_nsec=cn, capture_sec=ps, capture_nsec=pn, crawl_label=label)
    return ("tg" if magic == b"TGB1" else "yt"), cls(**arrays), cfg


def main(out):
    from distributed_crawler_b200.corpus import Corpus, YtCorpus
    os.makedirs(out, exist_ok=True)
    for name, n, profile, seed in CASES_TG:
        write_batch(os.path.join(out, name + ".tgb"), b"TGB1", Corpus(n, seed=seed, profile=profile, nthreads=1).batch)
    for name, n, seed in CASES_YT:
        write_batch(os.path.join(out, name + ".ytb"), b"YTB1", YtCorpus(n, seed=seed, nthreads=1).batch)
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "go_fixtures"))
