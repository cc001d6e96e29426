#!/bin/bash
echo "== chain tests"; timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_large.py -x -q -m gpu 2>&1 | tail -3
echo "== bench c3"; timeout 600 python bench.py --steps 10 --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-170
echo "== bench 8192 ch (dual)"; timeout 600 python bench.py --config c5 --channels 8192 --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | cut -c1-170
MSDR_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -18 | head -9
