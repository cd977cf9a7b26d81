#!/bin/bash
# on the GPU box: the config-2 bench step (10 M messages, resident) with the variants of build_variants/ and grid multipliers
cd "$(dirname "$0")/.."
cp distributed_crawler_b200/libtgingest.so /tmp/keep.so
for v in ${VARIANTS:-all7}; do
  cp build_variants/libtgingest_$v.so distributed_crawler_b200/libtgingest.so
  for m in ${MULTS:-48}; do
    for lm in ${LANE_MULTS:-3}; do
      echo "== $v TGI_GRID_MULT=$m TGI_LANE_MULT=$lm"
      TGI_LANE_MULT=$lm TGI_GRID_MULT=$m python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['roofline']['passes_ms'], round(d['roofline']['kernel_ms'],2))"
    done
  done
done
cp /tmp/keep.so distributed_crawler_b200/libtgingest.so
