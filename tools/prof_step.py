import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
c = Corpus(n, profile=2)
e = Engine()
e.telegram_upload(0, c.batch)
for i in range(3):
    r = e.telegram_run_resident(0, abi.RUN_JSONL | abi.RUN_NO_D2H)
    print(i, "kernel_ms", r.kernel_ms, "parse", r.parse_ms, "emit", r.emit_ms, "fixed", r.emit_main_ms, "bytes", r.jsonl_len, "GB/s emit", r.jsonl_len / r.emit_ms / 1e6)
