#!/bin/bash
# on the GPU box: one YouTube batch (config-4 shape, 2 M records) with the LB_YT variants of build_variants/
cd "$(dirname "$0")/.."
cp distributed_crawler_b200/libtgingest.so /tmp/keep.so
for v in ${VARIANTS:-yt3 yt4 yt5 yt6}; do
  cp build_variants/libtgingest_$v.so distributed_crawler_b200/libtgingest.so
  for m in ${MULTS:-128}; do
    echo "== $v TGI_GRID_MULT=$m"
    TGI_GRID_MULT=$m python - <<'PY'
import sys
sys.path.insert(0, ".")
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import YtCorpus
from distributed_crawler_b200.engine import Engine
c = YtCorpus(2_000_000, seed=0x5EED0004, nthreads=16)
e = Engine(frontier_capacity=1 << 23)
for i in range(3):
    e.frontier_clear()
    r = e.youtube(c.batch, abi.RUN_JSONL | abi.RUN_LINKS | abi.RUN_FRONTIER, copy=False)
print("kernel_ms", round(r.kernel_ms, 2), "parse+size", round(r.parse_ms, 2), "emit", round(r.emit_ms, 2), "Mrec/s", round(2.0 / r.kernel_ms * 1e3, 1))
PY
  done
done
cp /tmp/keep.so distributed_crawler_b200/libtgingest.so
