"""Phase clock of the page kernel (csrc/tg_page.cuh): TGI_PAGE_TRACE makes the library print the %globaltimer deltas between
the grid barriers of each call (stderr).  Third call of every size = warm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TGI_PAGE_TRACE"] = "1"
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
e = Engine()
flags = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
for n in (10, 100, 1000, 4000):
    c = Corpus(n, profile=2, nthreads=8)
    for _ in range(3):
        e.telegram(c.batch, flags, copy=False)
