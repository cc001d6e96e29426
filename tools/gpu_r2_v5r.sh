#!/bin/bash
for i in 1 2; do
echo -n "c5 full variant 1 sym: "; timeout 600 python bench.py --config c5 --variant 1 --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3))"
echo -n "c5 full variant 1 nosym: "; MSDR_NOSYM=1 timeout 600 python bench.py --config c5 --variant 1 --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3))"
done
