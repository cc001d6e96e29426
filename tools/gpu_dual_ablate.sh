#!/bin/bash
# ablation at 65536 channels x 128 blocks (two chain sets per SM)
for v in 0 16 32 48 512; do
  echo -n "65536 channels, variant $v: "
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --e2e-steps 0 --channels 65536 --seconds 0.3715 --blocks-per-update 128 --variant $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s')"
done 2>&1 | tee gpurun_out/dual_ablation.txt
