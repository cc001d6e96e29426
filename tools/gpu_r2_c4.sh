#!/bin/bash
mkdir -p gpurun_out/r2
echo "== 256-tap tests"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "long_taps" 2>&1 | tail -3
echo "== c4 at size"; timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -k "c4_at_size" 2>&1 | tail -3
echo "== bench c4 (default = v5l)"; timeout 900 python bench.py --config c4 --steps 5 --no-cpu --e2e-steps 0 2>gpurun_out/r2/c4_v5l.err | cut -c1-200; tail -2 gpurun_out/r2/c4_v5l.err
echo "== bench c4 (v4, variant 8192)"; timeout 900 python bench.py --config c4 --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant 8192 2>/dev/null | cut -c1-200
echo "== prof"; MSDR_PROF=1 timeout 600 python bench.py --config c4 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -8
