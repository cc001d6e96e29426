#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
MSDR_VARIANT=64 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_tc.sh 2>&1 | tail -6
bash tools/gpu_v4_ablate.sh
for v in 0; do
  echo "== variant $v"
  MSDR_PROF=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --seconds 0.4 --variant $v 2>&1 | grep -A8 "msdr prof" | head -9
done 2>&1 | tee gpurun_out/v4_prof.txt
