"""Per-call cost of tgi_telegram_batch for the batch sizes the reference's call sites produce (one page of one channel's
messages per call, crawl/runner.go:1110-1400): host buffers in, host JSONL out, blocking call.  Page-sized batches take the
one-launch path (csrc/tg_page.cuh); `pipeline` is the same call with TGI_NO_PAGE=1 (the ordinary multi-launch pipeline).
The loop calls the C entry point directly (ctypes, descriptor built once): the wrapper's Result object is not in the timing."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine, lib
e = Engine()
L = lib()
flags = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF


def timed(d, reps):
    r = abi.ResultC()
    for _ in range(3):
        assert L.tgi_telegram_batch(e.h, C.byref(d), flags, C.byref(r)) == 0
        L.tgi_result_release(e.h, r.slot)
    best, t0 = 1e9, time.perf_counter()
    for _ in range(reps):
        t = time.perf_counter()
        L.tgi_telegram_batch(e.h, C.byref(d), flags, C.byref(r))
        best = min(best, time.perf_counter() - t)
        L.tgi_result_release(e.h, r.slot)
    return (time.perf_counter() - t0) / reps, best, r


for n in (10, 100, 1000, 4000, 10000, 100000, 1000000):
    c = Corpus(n, profile=2, nthreads=8)
    d = c.batch.descriptor()
    reps = 200 if n <= 10000 else 5
    mean, best, r = timed(d, reps)
    line = f"n={n:8d}  call {mean*1e3:8.3f} ms (best {best*1e3:7.3f})  kernels {r.kernel_ms:7.3f} ms  {n/mean/1e6:8.2f} M msg/s  launches {r.gpu_launches:3d}"
    if r.gpu_launches == 1:
        os.environ["TGI_NO_PAGE"] = "1"
        m2, b2, r2 = timed(d, reps)
        del os.environ["TGI_NO_PAGE"]
        line += f"   | pipeline: call {m2*1e3:8.3f} ms (best {b2*1e3:7.3f})  launches {r2.gpu_launches}"
    print(line, flush=True)
