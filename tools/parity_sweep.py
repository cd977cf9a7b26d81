"""Extended randomized parity sweep (GPU vs oracle), beyond what tests/ runs every time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
from oracle.pyoracle import Oracle
from helpers import assert_results_equal
from yt_corpus import make_youtube, make_youtube_config4
from gm_corpus import make_generic

ALL = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
cases = 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    n = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 1000, 4097, 30000]))
    profile = int(rng.integers(1, 4))
    seed = int(rng.integers(1, 1 << 40))
    first = int(rng.integers(0, 1 << 30))
    cfg = dict(tz_offset_sec=int(rng.choice([0, 3600, -12600, 19800])), crawl_label=bytes(rng.choice([b"", b"lbl", b'x"<y>&z\\'])),
               created_at_nsec=int(rng.choice([0, 5, 123000000])), capture_nsec=int(rng.choice([0, 999999999])))
    if rng.random() < 0.3:
        cfg["min_post_date"] = 1_700_000_000
    c = Corpus(n, seed=seed, first=first, profile=profile, nthreads=4)
    flags = int(rng.choice([ALL, abi.RUN_JSONL, abi.RUN_LINKS, abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_FILTER, ALL | abi.RUN_FILTER]))
    o, e = Oracle(**cfg), Engine(**cfg)
    ro, rg = o.telegram(c.batch, flags), e.telegram(c.batch, flags)
    assert_results_equal(ro, rg, flags, f"tg n={n} profile={profile} seed={seed} flags={flags}")
    if flags & abi.RUN_FRONTIER:
        assert np.array_equal(o.frontier_export(), e.frontier_export())
    e.close()
    cases += 1
for seed in range(6):
    for mk, n in ((make_youtube, 700), (make_youtube_config4, 1500)):
        b, _, _ = mk(n, seed=100 + seed)
        o, e = Oracle(), Engine()
        f = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER
        assert_results_equal(o.youtube(b, f), e.youtube(b, f), f, f"yt {mk.__name__} seed={seed}")
        assert np.array_equal(o.frontier_export(), e.frontier_export())
        e.close()
        cases += 1
    gb, _ = make_generic(900, seed=200 + seed)
    o, e = Oracle(), Engine()
    assert_results_equal(o.generic(gb), e.generic(gb), abi.RUN_JSONL, f"generic seed={seed}")
    e.close()
    cases += 1
print("parity sweep ok:", cases, "cases")
