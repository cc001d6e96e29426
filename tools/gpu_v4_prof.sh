#!/bin/bash
# per-role cycle profile of the tensor-core chain kernel on the C3 shape (first update only), full and ablated
mkdir -p gpurun_out
for v in ${VARIANTS:-0 16 32 48}; do
  echo "== variant $v"
  MSDR_PROF=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --seconds 0.4 --variant $v 2>&1 | grep -A8 "msdr prof" | head -9
done 2>&1 | tee gpurun_out/v4_prof.txt
