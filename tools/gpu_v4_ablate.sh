#!/bin/bash
# ablation of the tensor-core chain kernel: ms per 1.8 G-sample step with parts of the arithmetic switched off (results wrong)
mkdir -p gpurun_out
for v in 0 16 32 48 256 128 64; do
  echo -n "variant $v: "
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --variant $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'ms/step', round(d['value']), 'Msamples/s', d['roofline']['avg_launch_ms'])"
done 2>&1 | tee gpurun_out/v4_ablation.txt
