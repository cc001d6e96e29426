#!/bin/bash
for v in 0 16384; do
echo "== variant $v one launch per step"; timeout 600 python bench.py --variant $v --steps 10 --no-cpu --no-parity --e2e-steps 0 --seconds 2.97 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d.get('gpu_launches'))"
done
