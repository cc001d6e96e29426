#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
echo "== v5 forced small"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "every_kernel_shape" 2>&1 | tail -3
echo "== at size"; timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -k "row_block or c5_interleaved" 2>&1 | tail -3
for v in 0 1 2; do for ch in 131072; do
  echo "== bench c5 $ch ch variant $v"; timeout 600 python bench.py --config c5 --channels $ch --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | cut -c1-170
done; done
echo "== prof 131072 v0"; MSDR_PROF=1 timeout 600 python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --variant 0 2>&1 >/dev/null | tail -8 | tee $O/v5e_prof.txt
for v in 16 32 48; do echo "== ablate $v"; timeout 600 python bench.py --config c5 --channels 131072 --steps 5 --no-cpu --no-parity --e2e-steps 0 --variant $v 2>/dev/null | cut -c1-160; done
B="python bench.py --config c5 --channels 131072 --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 0"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -f -o gpurun_out/r2/prof_v5e $B > gpurun_out/r2/prof_v5e.log 2>&1
ls -la gpurun_out/r2/prof_v5e.ncu-rep
