#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, pipe microbenchmarks, bench line.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
echo "== pipes"; timeout 300 tools/microbench/pipes > gpurun_out/pipes.txt 2>&1; tail -25 gpurun_out/pipes.txt
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
