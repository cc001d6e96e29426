#!/bin/bash
echo "== chain + large tests"; timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_large.py -x -q -m gpu 2>&1 | tail -3
for v in 0 0 16384; do
echo -n "variant $v: "; timeout 600 python bench.py --variant $v --steps 10 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['roofline']['kernel'][:12])"
done
