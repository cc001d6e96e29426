#!/bin/bash
for v in 0 0; do
echo -n "variant $v: "; timeout 600 python bench.py --variant $v --steps 10 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['roofline']['kernel'][:12])"
done
MSDR_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 --seconds 2.97 2>&1 >/dev/null | tail -11 | cut -c1-150
