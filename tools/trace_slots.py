import os, sys, time
sys.path.insert(0, "/root/repo")
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
CH = 500_000
e = Engine(frontier_capacity=1 << 24)
F = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
if len(sys.argv) > 1: F |= abi.RUN_NO_D2H
c = Corpus(CH * 3, seed=0x5EED0002, profile=2, nthreads=32)
staged = [e.stage(c.batch.slice(k * CH, (k + 1) * CH)) for k in range(3)]
inflight = []
for i in range(12):
    k = i % 3
    if len(inflight) == 3:
        s0 = inflight.pop(0); e.telegram_wait(s0); e.release(s0)
    if i == 6: sys.stderr.write("---- steady state\n")
    e.telegram_submit(k, staged[k], F)
    inflight.append(k)
for s0 in inflight:
    e.telegram_wait(s0); e.release(s0)
