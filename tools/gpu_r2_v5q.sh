#!/bin/bash
echo "== v5 tests"; timeout 900 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape and 409 or row_block" 2>&1 | tail -2
for i in 1 2; do
echo -n "c5 131072 sym: "; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3))"
echo -n "c5 131072 nosym: "; MSDR_NOSYM=1 timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3))"
done
echo -n "c5 full: "; timeout 600 python bench.py --config c5 --steps 5 --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['parity_checked'])"
