#!/bin/bash
# headline step time vs the streamed-weight n-tile width of the pointwise tensor-core GEMM (AC_TC_BN_STREAM)
for bn in 128 96 64 48 32; do
  AC_TC_BN_STREAM=$bn timeout 300 python scripts/effb2_chunk_sweep.py 0,0 2>/dev/null | grep chunk | sed "s/^/bn=$bn /"
done
