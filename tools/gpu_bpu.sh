#!/bin/bash
for bpu in 64 128 256 1024 3446; do for v in 0 128; do
  echo -n "blocks/update $bpu variant $v: "
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --variant $v --blocks-per-update $bpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'ms/step', round(d['value']), 'Msamples/s', d['roofline']['avg_launch_ms'])"
done; done 2>&1 | tee gpurun_out/bpu.txt
