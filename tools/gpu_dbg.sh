#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_chain.py -q -k "every_kernel_shape" 2>&1 | grep -v "^$" | tail -40
for v in 0 256 128 64; do
  echo -n "variant $v checksum: "
  timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --variant $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['result_checksum'], round(d['value']))"
done
