#!/bin/bash
for v in 0 32768; do for ch in 65536 131072 1048576; do
  echo "== bench c5 $ch variant $v"; timeout 600 python bench.py --config c5 --channels $ch --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | cut -c1-170
done; done
echo "== shapes on half-tile form"; MSDR_VARIANT=$((4096+32768)) timeout 900 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "not every_kernel_shape and not whole_file and not errors_match" 2>&1 | tail -2
