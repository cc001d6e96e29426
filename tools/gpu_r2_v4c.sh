#!/bin/bash
echo "== v4 tests"; timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_tc.py -x -q -m gpu -k "every_kernel_shape or long_taps or tc" 2>&1 | tail -2
for c in 4096 8192 11264; do
echo -n "channels $c v4: "; timeout 600 python bench.py --channels $c --seconds 2.97 --variant 16384 --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['roofline']['kernel'][:60])"
done
