#!/bin/bash
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape and default or window_sizes or ragged" 2>&1 | tail -2
for i in 1 2 3; do
timeout 600 python bench.py --steps 10 --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['parity_checked']['device_resident']['mismatches'])"
done
