#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
echo "== shapes"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape" 2>&1 | tail -2
echo "== at size"; timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -k "row_block or c5_interleaved" 2>&1 | tail -2
for v in 0 1 4; do for ch in 131072; do
  echo "== bench c5 $ch ch variant $v"; timeout 600 python bench.py --config c5 --channels $ch --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | cut -c1-170
done; done
echo "== bench c5 65536"; timeout 600 python bench.py --config c5 --channels 65536 --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | cut -c1-170
for v in 16 32 48; do echo "== ablate $v"; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | cut -c1-160; done
echo "== prof 131072 v0"; MSDR_PROF=1 timeout 600 python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --variant 0 2>&1 >/dev/null | tail -8 | tee $O/v5j_prof.txt
