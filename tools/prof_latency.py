"""Per-call cost of tgi_telegram_batch for the batch sizes the reference's call sites produce (one channel's messages per
call, crawl/runner.go:1110-1400): host buffers in, host JSONL out, blocking call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
e = Engine()
flags = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
for n in (100, 1000, 10000, 100000, 1000000):
    c = Corpus(n, profile=2, nthreads=8)
    for _ in range(3):
        e.telegram(c.batch, flags, copy=False)
    reps = 20 if n <= 10000 else 5
    t = time.perf_counter()
    for _ in range(reps):
        r = e.telegram(c.batch, flags, copy=False)
    dt = (time.perf_counter() - t) / reps
    print(f"n={n:8d}  call {dt*1e3:8.3f} ms  kernels {r.kernel_ms:7.3f} ms  {n/dt/1e6:8.2f} M msg/s  launches {r.gpu_launches}")
