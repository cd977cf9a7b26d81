"""BASELINE configs 3/5 shape: link extraction + frontier dedup only (no JSONL), one resident batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus, PROFILE_LINKS
from distributed_crawler_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
c = Corpus(n, profile=PROFILE_LINKS)
e = Engine()
e.telegram_upload(0, c.batch)
for i in range(4):
    e.frontier_clear()
    r = e.telegram_run_resident(0, abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF | abi.RUN_NO_D2H)
    print(i, "kernel_ms", round(r.kernel_ms, 3), "parse", round(r.parse_ms, 3), "links", r.n_links, "unique", r.frontier_size,
          "M msg/s", round(n / r.kernel_ms / 1e3, 1), "input GB/s", round(c.batch.input_bytes() / r.kernel_ms / 1e6, 1))
