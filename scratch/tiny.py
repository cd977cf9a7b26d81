import sys
sys.path.insert(0, '.')
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
c = Corpus(64, profile=2, nthreads=1)
e = Engine()
r = e.telegram(c.batch, abi.RUN_JSONL | abi.RUN_LINKS)
print("ok", r.jsonl_len)
