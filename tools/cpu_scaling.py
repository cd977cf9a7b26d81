"""Thread-scaling curve of the CPU arm (the oracle) on this host: msgs/s for 1..N threads, plus what the box says
about its CPUs (cgroup quota, sockets / NUMA), so that the cpu_baseline of bench.py can be read properly."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from oracle import pyoracle
from oracle.pyoracle import Oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpuset.cpus.effective"):
    try:
        print(f, open(f).read().strip())
    except OSError as e:
        print(f, e)
print(subprocess.run("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout)
print("sched_getaffinity", len(os.sched_getaffinity(0)))
c = Corpus(n, seed=0x5EED0002, profile=2)
F = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
o = Oracle()
T = os.cpu_count() or 1
ts = [1, 2, 4, 8, 16, 32, 64, 128, 256]
for pin in (0x20000, 0):
    for t in [x for x in ts if x <= T]:
        sub = c.batch if t >= 8 else c.batch.slice(0, n // 8)
        for _ in range(2):
            pyoracle.lib().orc_frontier_clear(o.h)
            t0 = time.perf_counter()
            o.telegram(sub, F | 0x10000 | pin, nthreads=t, copy=False)
            dt = time.perf_counter() - t0
        print(f"pin={bool(pin)} threads={t:4d} {sub.n / dt / 1e3:10.0f} K msg/s  {sub.n / dt / t / 1e3:8.1f} K/thread", flush=True)
