"""Parity pinning by bytes the reference ITSELF produced: go_ref/ runs the real ParseMessage /
convertVideoToPost / encoding/json over the seeded batches of go_ref/dump_batch.py and writes
tests/golden/go_fixtures/<name>.jsonl.  The reference cannot run in this image (no Go toolchain), so until a maintainer
commits those files the JSON bytes of the oracle stay "parity unpinned" and these tests skip with that reason.
What always runs: the batch file format round trip (the Go harness, the oracle and the CUDA path read identical inputs)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from distributed_crawler_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "go_fixtures")
spec = importlib.util.spec_from_file_location("dump_batch", os.path.join(ROOT, "go_ref", "dump_batch.py"))
dump_batch = importlib.util.module_from_spec(spec)
spec.loader.exec_module(dump_batch)


def _cases():
    return sorted(glob.glob(os.path.join(FIX, "*.jsonl")))


def _batch_for(jsonl):
    base = jsonl[: -len(".jsonl")]
    path = base + (".tgb" if os.path.exists(base + ".tgb") else ".ytb")
    return dump_batch.read_batch(path)


def test_batch_file_round_trip(tmp_path):
    from distributed_crawler_b200.corpus import Corpus, YtCorpus
    from oracle.pyoracle import Oracle
    c = Corpus(3000, seed=0x5EED0002, profile=2, nthreads=1)
    p = str(tmp_path / "a.tgb")
    dump_batch.write_batch(p, b"TGB1", c.batch)
    kind, b2, cfg = dump_batch.read_batch(p)
    assert kind == "tg" and all(np.array_equal(getattr(c.batch, k), getattr(b2, k)) for k in c.batch.FIELDS)
    flags = abi.RUN_JSONL | abi.RUN_LINKS
    assert np.array_equal(Oracle(**cfg).telegram(c.batch, flags).jsonl, Oracle(**cfg).telegram(b2, flags).jsonl)
    y = YtCorpus(500, nthreads=1)
    p = str(tmp_path / "a.ytb")
    dump_batch.write_batch(p, b"YTB1", y.batch)
    kind, y2, _ = dump_batch.read_batch(p)
    assert kind == "yt" and all(np.array_equal(getattr(y.batch, k), getattr(y2, k)) for k in y.batch.FIELDS)


def _expected_lines(result):
    return [result.line(i) for i in range(result.n) if result.status[i] == abi.ST_EMITTED]


@pytest.mark.skipif(not _cases(), reason="parity unpinned: no reference-produced fixtures under tests/golden/go_fixtures "
                                         "(go_ref/make_fixtures.sh needs Go + the reference checkout)")
def test_oracle_matches_reference_fixtures():
    from oracle.pyoracle import Oracle
    for jsonl in _cases():
        kind, batch, cfg = _batch_for(jsonl)
        o = Oracle(**cfg)
        r = (o.youtube if kind == "yt" else o.telegram)(batch, abi.RUN_JSONL | abi.RUN_LINKS)
        want = open(jsonl, "rb").read().splitlines(keepends=True)
        got = _expected_lines(r)
        assert len(got) == len(want), f"{jsonl}: {len(got)} lines vs {len(want)} from the reference"
        for i, (a, b) in enumerate(zip(got, want)):
            assert a == b, f"{jsonl}: line {i} differs from the reference's bytes"
        links = os.path.splitext(jsonl)[0] + ".links.txt"
        if os.path.exists(links):
            for row in open(links, encoding="utf-8"):
                idx, _, names = row.rstrip("\n").partition("\t")
                mine = sorted(n.decode() for n, _ in r.record_links(int(idx)))
                assert mine == ([x for x in names.split(",") if x]), f"{jsonl}: outlinks of record {idx}"


@pytest.mark.gpu
@pytest.mark.skipif(not _cases(), reason="parity unpinned: no reference-produced fixtures under tests/golden/go_fixtures")
def test_gpu_matches_reference_fixtures():
    from distributed_crawler_b200.engine import Engine
    for jsonl in _cases():
        kind, batch, cfg = _batch_for(jsonl)
        e = Engine(**cfg)
        r = (e.youtube if kind == "yt" else e.telegram)(batch, abi.RUN_JSONL | abi.RUN_LINKS)
        assert b"".join(_expected_lines(r)) == open(jsonl, "rb").read(), jsonl
        e.close()
