#!/bin/bash
mkdir -p gpurun_out
echo "== shape test"; timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape and (default or 16384)" 2>&1 | tail -3
for v in 0 2; do
echo "== bench c3 variant $v"; timeout 600 python bench.py --variant $v --steps 5 --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-170
MSDR_PROF=1 MSDR_PROF_CTAS=1 timeout 600 python bench.py --variant $v --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 --seconds 2.97 > gpurun_out/v6_prof_$v.txt 2>&1
grep -A10 "prof v6" gpurun_out/v6_prof_$v.txt | tail -10 | cut -c1-150
done
