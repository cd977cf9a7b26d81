"""One YouTube batch (tests/yt_corpus.py shapes) through the engine: device time per pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from yt_corpus import make_youtube, make_youtube_config4
from distributed_crawler_b200 import abi
from distributed_crawler_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
t = time.time(); b, _, _ = (make_youtube(n, seed=21) if "--adversarial" in sys.argv else make_youtube_config4(n)); print("corpus", round(time.time() - t, 1), "s", flush=True)
e = Engine()
for i in range(3):
    r = e.youtube(b, abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER, copy=False)
    print(i, "kernel_ms", round(r.kernel_ms, 3), "parse+size", round(r.parse_ms, 3), "emit", round(r.emit_ms, 3), "bytes", r.jsonl_len,
          "Mrec/s", round(n / r.kernel_ms / 1e3, 2), "GB/s out", round(r.jsonl_len / r.kernel_ms / 1e6, 1))
if "--oracle" in sys.argv:
    from oracle.pyoracle import Oracle
    o = Oracle()
    nt = os.cpu_count() or 1
    o.youtube(b, abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER, nthreads=nt, copy=False)
    t = time.time(); o.youtube(b, abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER, nthreads=nt, copy=False); dt = time.time() - t
    print("oracle (C restatement, %d threads, warm): %.2f Mrec/s" % (nt, n / dt / 1e6))
