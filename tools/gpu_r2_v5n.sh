#!/bin/bash
echo "== v5 tests"; timeout 900 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape or long_taps or row_block" 2>&1 | tail -2
echo "== c5 131072"; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | cut -c1-170
echo "== c5 full"; timeout 600 python bench.py --config c5 --steps 5 --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-170
echo "== c4"; timeout 600 python bench.py --config c4 --steps 5 --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-170
