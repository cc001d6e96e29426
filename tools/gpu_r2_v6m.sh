#!/bin/bash
for lib in "" minimal-sdr_b200/csrc/variants/libmsdr_lazy01.so "" minimal-sdr_b200/csrc/variants/libmsdr_lazy01.so; do
echo -n "lib '$lib': "; MSDR_LIBMSDR=$lib timeout 600 python bench.py --steps 10 --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['parity_checked']['device_resident']['mismatches'])"
done
