#!/bin/bash
# C3 against the update length: time-folded kernel (0) and chain kernel (16384)
mkdir -p gpurun_out/r2v6
( for b in 8 32 64 256 1024; do for v in 0 16384; do
  echo -n "blocks per update $b variant $v: "
  timeout 300 python bench.py --blocks-per-update $b --seconds 2.97 --steps 3 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s', d['gpu_launches'], 'launches', d['roofline']['kernel'][:14])"
done; done ) 2>&1 | tee gpurun_out/r2v6/bpu.txt
