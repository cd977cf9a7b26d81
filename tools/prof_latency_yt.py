"""Per-call cost of tgi_youtube_batch for Data-API-page-sized batches (50 videos per page): page path vs bulk pipeline."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from distributed_crawler_b200 import abi
from distributed_crawler_b200.engine import Engine, lib
from yt_corpus import make_youtube_config4
e = Engine()
L = lib()
flags = abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER


def timed(d, reps):
    r = abi.ResultC()
    for _ in range(3):
        assert L.tgi_youtube_batch(e.h, C.byref(d), flags, C.byref(r)) == 0
        L.tgi_result_release(e.h, r.slot)
    best, t0 = 1e9, time.perf_counter()
    for _ in range(reps):
        t = time.perf_counter()
        L.tgi_youtube_batch(e.h, C.byref(d), flags, C.byref(r))
        best = min(best, time.perf_counter() - t)
        L.tgi_result_release(e.h, r.slot)
    return (time.perf_counter() - t0) / reps, best, r


for n in (10, 50, 500, 2000):
    batch, _, _ = make_youtube_config4(n, seed=n)
    d = batch.descriptor()
    mean, best, r = timed(d, 200)
    line = f"n={n:6d}  call {mean*1e3:8.3f} ms (best {best*1e3:7.3f})  kernels {r.kernel_ms:7.3f} ms  launches {r.gpu_launches:3d}  {r.jsonl_len} JSONL bytes"
    if r.gpu_launches == 1:
        os.environ["TGI_NO_PAGE"] = "1"
        m2, b2, r2 = timed(d, 200)
        del os.environ["TGI_NO_PAGE"]
        line += f"   | pipeline: call {m2*1e3:8.3f} ms (best {b2*1e3:7.3f})  launches {r2.gpu_launches}"
    print(line, flush=True)
