"""Every product path once on small inputs: run under `compute-sanitizer --tool memcheck` / `--tool racecheck`."""
import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
from yt_corpus import make_youtube, make_youtube_config4
from gm_corpus import make_generic
e = Engine()
f = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
for p in (1, 2, 3):
    c = Corpus(3000, profile=p, nthreads=4)
    r = e.telegram(c.batch, f)
    print("tg", p, r.jsonl_len, r.n_links)
for mk in (make_youtube, make_youtube_config4):
    b, _, _ = mk(600, seed=3)
    r = e.youtube(b, abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER)
    print("yt", r.jsonl_len, r.n_links)
g, _ = make_generic(500, seed=4)
print("gm", e.generic(g).jsonl_len)
print("join", e.key_join(np.arange(2000).reshape(-1, 2), np.arange(1000).reshape(-1, 2))[:4])
print("done")
