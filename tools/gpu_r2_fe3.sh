#!/bin/bash
echo "== frontend tests"; timeout 600 python -m pytest tests/test_gpu_frontend.py -x -q -m gpu 2>&1 | tail -2
echo -n "c3 with frontend: "; timeout 600 python bench.py --with-frontend --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3))"
echo -n "c3 with frontend serial: "; timeout 600 python bench.py --with-frontend --frontend-serial --steps 5 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3))"
