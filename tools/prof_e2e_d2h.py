"""Where the e2e leg's device->host rate goes: the 3-slot pipeline of bench.py's e2e leg, the same with the input already
resident (no H2D in the loop; three threads, one slot each, blocking tgi_telegram_run_resident), and with the H2D but no
result copy — against tools/pcie_peak.py's 54 GB/s (one direction alone) / 45 GB/s (both directions busy)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
CH, ROUNDS = 500_000, 6
e = Engine(frontier_capacity=1 << 24)
F = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF
c = Corpus(CH * 3, seed=0x5EED0002, profile=2, nthreads=32)
staged = [e.stage(c.batch.slice(k * CH, (k + 1) * CH)) for k in range(3)]


def pipeline(flags, rounds):
    e.frontier_clear()
    t = time.perf_counter()
    d2h, inflight = 0, []
    for i in range(rounds * 3):
        k = i % 3
        if len(inflight) == 3:
            s0 = inflight.pop(0)
            d2h += e.telegram_wait(s0).d2h_bytes()
            e.release(s0)
        e.telegram_submit(k, staged[k], flags)
        inflight.append(k)
    for s0 in inflight:
        d2h += e.telegram_wait(s0).d2h_bytes()
        e.release(s0)
    dt = time.perf_counter() - t
    return d2h / dt / 1e9, rounds * 3 * CH / dt / 1e6


def resident(flags, rounds):
    for k in range(3):
        e.telegram_upload(k, staged[k])
    e.frontier_clear()
    out = [0, 0, 0]

    def work(k):
        for _ in range(rounds):
            out[k] += e.telegram_run_resident(k, flags).d2h_bytes()

    th = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    t = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t
    return sum(out) / dt / 1e9, rounds * 3 * CH / dt / 1e6


for name, fn, flags in (("H2D + kernels + D2H (the e2e leg)", pipeline, F), ("kernels + D2H, input resident", resident, F),
                        ("H2D + kernels, no result copy", pipeline, F | abi.RUN_NO_D2H)):
    fn(flags, 2)
    g, m = fn(flags, ROUNDS)
    print(f"{name:40s} {g:6.1f} GB/s out   {m:6.1f} M msg/s", flush=True)

# the upload alone: one slot, then three slots from three threads (tgi_telegram_upload returns when the copy has landed)
in_bytes = sum(s.input_bytes() for s in staged) / 3
t = time.perf_counter()
for _ in range(6):
    e.telegram_upload(0, staged[0])
dt = time.perf_counter() - t
print(f"upload alone, one slot                   {6 * in_bytes / dt / 1e9:6.1f} GB/s in", flush=True)


def up(k):
    for _ in range(6):
        e.telegram_upload(k, staged[k])


th = [threading.Thread(target=up, args=(k,)) for k in range(3)]
t = time.perf_counter()
for x in th: x.start()
for x in th: x.join()
dt = time.perf_counter() - t
print(f"upload alone, three slots                {18 * in_bytes / dt / 1e9:6.1f} GB/s in", flush=True)
# kernels alone on a resident batch
for k in range(3):
    r = e.telegram_run_resident(k, F | abi.RUN_NO_D2H)
t = time.perf_counter()
for _ in range(6):
    r = e.telegram_run_resident(0, F | abi.RUN_NO_D2H)
dt = time.perf_counter() - t
print(f"kernels alone (resident, no copies)      {dt / 6 * 1e3:6.2f} ms per {CH} records, kernel_ms {r.kernel_ms:.2f}", flush=True)
