#!/bin/bash
# Debug: build the library with different lanes-per-query settings of the chamfer search and time tests/tools/bench_nnd.py
cd "$(dirname "$0")/.."
for ql in 1 2 4 8; do
  out=self6dpp_b200/lib/libdibr_b200_ql$ql.so
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -shared -cudart static -DDIBR_NND_QL=$ql -o $out self6dpp_b200/csrc/*.cu 2>/dev/null
  echo "QL=$ql"; DIBR_B200_LIB=$PWD/$out python tests/tools/bench_nnd.py 2>&1 | tail -1 | cut -c1-180
  rm -f $out
done
