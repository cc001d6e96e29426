#!/bin/bash
MSDR_PROF=1 timeout 600 python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -8
for v in 16 32 48; do echo -n "variant $v: "; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"; done
