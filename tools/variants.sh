#!/bin/bash
# Build libtgingest.so variants with different __launch_bounds__ budgets (LB_*: resident CTAs per SM the compiler budgets
# registers for) into build_variants/; tools/variants_bench.sh (Telegram, config-2 step) and tools/variants_yt.sh (YouTube) time them on the GPU box.  Usage: tools/variants.sh
set -e
cd "$(dirname "$0")/../distributed_crawler_b200/csrc"
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall -shared"
mkdir -p ../../build_variants
build() { name=$1; shift; $NV "$@" -o ../../build_variants/libtgingest_$name.so tgingest.cu -lcudart 2>/dev/null & }
for k in ${LBS:-4 5 6 7 8}; do
  build all$k -DLB_SIZE=$k -DLB_ESC=$k -DLB_MAPS=$k -DLB_PARSE=$k
  build yt$k -DLB_YT=$k
done
wait
ls -la ../../build_variants
