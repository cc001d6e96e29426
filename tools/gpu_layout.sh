#!/bin/bash
for lay in rows updates; do for v in 0 64 48; do
  echo -n "layout $lay variant $v: "
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --variant $v --layout $lay 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'ms/step', round(d['value']), 'Msamples/s', d['roofline']['avg_launch_ms'])"
done; done 2>&1 | tee gpurun_out/layout.txt
