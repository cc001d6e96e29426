#!/bin/bash
mkdir -p gpurun_out/r2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  MSDR_VARIANT=4096 timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "not every_kernel_shape and not whole_file and not errors_match" 2>&1 | grep -E "passed|failed|FAILED|Error|mismatch" | head -5
done
echo "== bench c3"; timeout 900 python bench.py --steps 5 --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-200
echo "== bench c5 131072"; timeout 900 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --e2e-steps 0 --no-parity 2>/dev/null | cut -c1-200
