#!/bin/bash
# channel counts between the three fused kernels, C3 launch shape (1024 blocks per update)
mkdir -p gpurun_out/r2v6
echo "== large tests"; timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu 2>&1 | tail -2
( for c in 4736 5120 6144 8192 9472 9600 10240 11264 12288 16384; do for v in 0 8192 4096; do
  echo -n "channels $c variant $v: "
  timeout 300 python bench.py --channels $c --seconds 2.97 --steps 3 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'Msamples/s', d['roofline']['kernel'][:14])"
done; done ) 2>&1 | tee gpurun_out/r2v6/channels_mid.txt
