#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
echo "== chunks test on v5"; MSDR_VARIANT=4096 timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "host_update_pipeline" 2>&1 | grep -E "Error|assert|mismatch|passed|failed" | head -12
echo "== v5 whole file"; timeout 900 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "whole_file" 2>&1 | tail -5
echo "== at size"; timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu 2>&1 | tail -5
for ch in 65536 131072; do
  echo "== bench c5 $ch ch"; timeout 600 python bench.py --config c5 --channels $ch --steps 5 --no-cpu --e2e-steps 0 > $O/v5b_c5_$ch.json 2> $O/v5b_c5_$ch.err; cut -c1-220 $O/v5b_c5_$ch.json; tail -3 $O/v5b_c5_$ch.err
done
echo "== prof 131072"; MSDR_PROF=1 timeout 600 python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>&1 >/dev/null | tail -8 | tee $O/v5b_prof.txt
for v in 16 32 48; do echo "== ablate $v"; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | cut -c1-160; done
