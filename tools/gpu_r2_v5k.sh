#!/bin/bash
echo "== shapes"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape" 2>&1 | tail -2
for v in 0 8 24 16; do
  echo "== bench c5 131072 variant $v"; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | cut -c1-170
done
