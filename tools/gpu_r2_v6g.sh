#!/bin/bash
mkdir -p gpurun_out
for lay in updates rows; do for v in 0 16384; do
echo -n "layout $lay variant $v: "; timeout 600 python bench.py --layout $lay --variant $v --steps 10 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['roofline']['kernel'][:12])"
done; done
MSDR_PROF=1 MSDR_PROF_CTAS=1 timeout 600 python bench.py --layout rows --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 --seconds 3.5 2>&1 | grep -B140 "prof v6" | grep -A140 "prof v6" | head -140 > gpurun_out/v6_ctas.txt
head -11 gpurun_out/v6_ctas.txt | cut -c1-150
