#!/bin/bash
mkdir -p gpurun_out
for v in 0; do
MSDR_PROF=1 timeout 600 python bench.py --variant $v --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 --seconds 2.97 > gpurun_out/abl_$v.txt 2>&1
echo "== variant $v"; grep -A10 "prof v6" gpurun_out/abl_$v.txt | tail -10 | cut -c1-158
done
