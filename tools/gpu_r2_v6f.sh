#!/bin/bash
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
for i in 1 2; do
for v in 0 16384; do
echo "== variant $v"; timeout 600 python bench.py --variant $v --steps 10 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline']['kernel'][:12])"
done; done
MSDR_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 --seconds 2.97 2>&1 | grep -A10 "prof v6" | tail -10 | cut -c1-150
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
