#!/bin/bash
mkdir -p gpurun_out/r2
echo "== frontend tests"; timeout 600 python -m pytest tests/test_gpu_frontend.py -x -q -m gpu 2>&1 | tail -3
echo "== default"; timeout 600 python bench.py --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | cut -c1-170
echo "== with-frontend serial"; timeout 600 python bench.py --steps 5 --no-cpu --no-parity --e2e-steps 0 --with-frontend --frontend-serial 2>/dev/null | cut -c1-170
echo "== with-frontend pipelined"; timeout 600 python bench.py --steps 5 --no-cpu --no-parity --e2e-steps 0 --with-frontend 2>gpurun_out/r2/fe_pipe.err | cut -c1-170; tail -2 gpurun_out/r2/fe_pipe.err
