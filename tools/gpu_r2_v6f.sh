#!/bin/bash
for i in 1 2; do
echo -n "ALU epilogue (default): "; timeout 600 python bench.py --steps 10 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3))"
echo -n "IMAD epilogue: "; MSDR_LIBMSDR=minimal-sdr_b200/csrc/variants/libmsdr_epiimad.so timeout 600 python bench.py --steps 10 --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['parity_checked']['device_resident']['mismatches'])"
done
