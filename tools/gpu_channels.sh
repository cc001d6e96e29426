#!/bin/bash
# throughput against channel count (device-resident, 128 blocks per update)
for c in 4096 8192 16384 65536 262144; do for v in 0 512; do
  echo -n "channels $c variant $v: "
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --e2e-steps 0 --channels $c --seconds 0.3715 --blocks-per-update 128 --variant $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"
done; done 2>&1 | tee gpurun_out/channels.txt
