#!/bin/bash
mkdir -p gpurun_out
echo "== chain tests"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "not whole_file" 2>&1 | tail -3
for a in 0 16; do
echo "== bench c3 ablate $a"; MSDR_ABLATE=$a timeout 600 python bench.py --steps 5 --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-170
MSDR_ABLATE=$a MSDR_PROF=1 MSDR_PROF_CTAS=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 --seconds 2.97 > gpurun_out/v6_prof_$a.txt 2>&1
grep -A10 "prof v6" gpurun_out/v6_prof_$a.txt | tail -10 | cut -c1-150
done
