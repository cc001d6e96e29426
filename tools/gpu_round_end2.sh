#!/bin/bash
# round end, second part: the with-frontend line of the final build
mkdir -p gpurun_out
timeout 900 python bench.py --with-frontend --no-cpu --e2e-steps 0 > gpurun_out/bench_fe.json 2> gpurun_out/bench_fe.err; cut -c1-200 gpurun_out/bench_fe.json
