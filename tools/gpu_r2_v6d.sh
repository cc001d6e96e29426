#!/bin/bash
mkdir -p gpurun_out
for sec in 2.97; do
MSDR_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 --seconds $sec > gpurun_out/len_$sec.txt 2>&1
echo "== seconds $sec"; grep -A10 "prof v6" gpurun_out/len_$sec.txt | tail -10 | cut -c1-60 | sed -n '4,5p;10p'
done
