#!/bin/bash
echo "== shape test"; timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape" 2>&1 | tail -5
echo "== bench c3"; timeout 600 python bench.py --steps 10 --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-400
MSDR_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -10
echo "== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
