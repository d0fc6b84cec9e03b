#!/bin/bash
# A/B timing of the tcgen05 attention variants at the lang-config shapes (developer tool): builds
# tools/micro/att_bench.cu (which includes the kernel source) once per variant.
cd "$(dirname "$0")/micro"
for cfg in "1 3 1" "1 0 1" "1 3 0" "0 3 1" "0 0 1"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
       -DSS_ATT_WIDE=$1 -DSS_ATT_POLY=$2 -DSS_ATT_PP=$3 -o att_bench_v att_bench.cu || exit 1
  echo "== wide=$1 poly=$2 pingpong=$3"
  ./att_bench_v 299277 16 48 5; ./att_bench_v 120000 16 32 5; ./att_bench_v 299277 2 16 5; ./att_bench_v 34000 16 16 5
done
